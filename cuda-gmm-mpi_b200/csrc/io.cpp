// io.cpp — input readers and output writers with the reference's file
// formats (readData.cpp:25-129; gaussian.cu:998-1061, 1180-1201).  One-shot
// I/O: kept format-compatible, not accelerated.
#include <cerrno>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "host_math.h"
#ifdef _OPENMP
#include <omp.h>
#endif

extern "C" {

void gmm_free(void* p) { std::free(p); }

// "*.bin": int32 nevents, int32 ndims, float32[nevents*ndims]   (readData.cpp:35-47)
static float* read_bin(const char* path, int* ndims, int* nevents) {
    FILE* f = std::fopen(path, "rb");
    if (!f) { gmm::set_error(std::string("cannot open ") + path); return nullptr; }
    int32_t hdr[2];
    if (std::fread(hdr, sizeof(int32_t), 2, f) != 2 || hdr[0] <= 0 || hdr[1] <= 0) {
        std::fclose(f); gmm::set_error("bad .bin header"); return nullptr;
    }
    const size_t count = (size_t)hdr[0] * (size_t)hdr[1];
    float* data = (float*)std::malloc(sizeof(float) * count);
    if (!data) { std::fclose(f); gmm::set_error("out of memory reading events"); return nullptr; }
    const size_t got = std::fread(data, sizeof(float), count, f);
    std::fclose(f);
    if (got != count) { std::free(data); gmm::set_error("truncated .bin file"); return nullptr; }
    *nevents = hdr[0];
    *ndims = hdr[1];
    return data;
}

// Anything else: comma-separated text; the first non-empty line is a header
// and is dropped; the number of columns is the header's comma count + 1;
// every following non-empty line must provide that many fields, parsed as
// atof does; a short line makes the whole read fail (readData.cpp:49-129).
static float* read_csv(const char* path, int* ndims, int* nevents) {
    FILE* f = std::fopen(path, "r");
    if (!f) { gmm::set_error(std::string("Unable to read the file ") + path); return nullptr; }
    std::vector<std::string> lines;
    {
        std::string cur;
        int ch;
        while ((ch = std::fgetc(f)) != EOF) {
            if (ch == '\n') { if (!cur.empty()) lines.push_back(cur); cur.clear(); }
            else cur.push_back((char)ch);
        }
        if (!cur.empty()) lines.push_back(cur);
    }
    std::fclose(f);
    if (lines.empty()) { gmm::set_error("empty input file"); return nullptr; }
    int dims = 0;
    {   // strtok semantics: runs of ',' are one separator, empty tokens vanish
        const std::string& h = lines[0];
        size_t i = 0;
        while (i < h.size()) {
            while (i < h.size() && h[i] == ',') i++;
            if (i >= h.size()) break;
            dims++;
            while (i < h.size() && h[i] != ',') i++;
        }
    }
    const int n = (int)lines.size() - 1;
    if (dims <= 0 || n <= 0) { gmm::set_error("no data rows in input file"); return nullptr; }
    float* data = (float*)std::malloc(sizeof(float) * (size_t)dims * n);
    if (!data) { gmm::set_error("out of memory reading events"); return nullptr; }
    for (int r = 0; r < n; r++) {
        const std::string& s = lines[r + 1];
        size_t i = 0;
        for (int d = 0; d < dims; d++) {
            while (i < s.size() && s[i] == ',') i++;
            if (i >= s.size()) { std::free(data); gmm::set_error("inconsistent number of dimensions"); return nullptr; }
            size_t j = i;
            while (j < s.size() && s[j] != ',') j++;
            data[(size_t)r * dims + d] = (float)std::atof(s.substr(i, j - i).c_str());
            i = j;
        }
    }
    *ndims = dims;
    *nevents = n;
    return data;
}

float* gmm_read_data(const char* path, int* ndims, int* nevents) {
    if (!path || !ndims || !nevents) { gmm::set_error("gmm_read_data: bad argument"); return nullptr; }
    const size_t len = std::strlen(path);
    if (len >= 3 && std::strcmp(path + len - 3, "bin") == 0) return read_bin(path, ndims, nevents);   // readData.cpp:28
    return read_csv(path, ndims, nevents);
}

// writeCluster (gaussian.cu:1180-1197) for every saved cluster (:1024-1040).
int gmm_write_summary(const char* path, const clusters_t* c, int K, int D) {
    FILE* f = std::fopen(path, "w");
    if (!f) return gmm::fail(GMM_ERR_IO, std::string("Unable to open file '") + path + "' for writing.");
    for (int k = 0; c && k < K; k++) {
        std::fprintf(f, "Cluster #%d\n", k);
        std::fprintf(f, "Probability: %f\n", c->pi[k]);
        std::fprintf(f, "N: %f\n", c->N[k]);
        std::fprintf(f, "Means: ");
        for (int i = 0; i < D; i++) std::fprintf(f, "%.3f ", c->means[(size_t)k * D + i]);
        std::fprintf(f, "\n\nR Matrix:\n");
        for (int i = 0; i < D; i++) {
            for (int j = 0; j < D; j++) std::fprintf(f, "%.3f ", c->R[(size_t)k * D * D + i * D + j]);
            std::fprintf(f, "\n");
        }
        std::fprintf(f, "\n\n");
    }
    std::fclose(f);
    return GMM_OK;
}

// "%f" of a float, digit for digit what printf produces, without the per-call cost of fprintf (the .results file of
// 10M events x (24 + 64) columns is 880M numbers).  A float's fraction times 10^6 is exact in double (24 significant
// bits x 15625 x 2^6), so the round-half-even decision is exact too.
static inline char* fmt_f6(char* p, float v) {
    if (!std::isfinite(v) || std::fabs(v) >= 1.0e15f) return p + std::sprintf(p, "%f", v);
    double x = v;
    if (std::signbit(v)) { *p++ = '-'; x = -x; }
    unsigned long long ip = (unsigned long long)x;
    const double f6 = (x - (double)ip) * 1.0e6;
    unsigned long long r = (unsigned long long)f6;
    const double rem = f6 - (double)r;
    if (rem > 0.5 || (rem == 0.5 && (r & 1ull))) r++;
    if (r == 1000000ull) { r = 0; ip++; }
    char tmp[24];
    int n = 0;
    do { tmp[n++] = (char)('0' + ip % 10); ip /= 10; } while (ip);
    while (n) *p++ = tmp[--n];
    *p++ = '.';
    for (int i = 5; i >= 0; i--) { p[i] = (char)('0' + r % 10); r /= 10; }
    return p + 6;
}

// .results (gaussian.cu:1042-1059): "x1,...,xD<TAB>g1,...,gK\n", all %f;
// memberships are cluster-major [K][N].  Rows are formatted in parallel into per-block buffers and written in order.
int gmm_write_results(const char* path, const float* ev, long long N, int D, const clusters_t* c, int K) {
    FILE* f = std::fopen(path, "w");
    if (!f) return gmm::fail(GMM_ERR_IO, std::string("Unable to open file '") + path + "' for writing.");
    const long long BR = 1024;                                   // rows per block
    const long long nblocks = (N + BR - 1) / BR;
    int nt = 1;
#ifdef _OPENMP
    nt = omp_get_max_threads();
    if (nt > 32) nt = 32;
#endif
    const long long group = (long long)nt * 2;
    std::vector<std::vector<char>> bufs((size_t)group);
    std::vector<size_t> used((size_t)group, 0);
    const size_t cap = (size_t)BR * ((size_t)(D + K) * 49 + 2);     // "%f" of FLT_MAX is 46 characters
    bool ok = true;
    for (long long g0 = 0; g0 < nblocks && ok; g0 += group) {
        const long long g1 = g0 + group < nblocks ? g0 + group : nblocks;
#pragma omp parallel for schedule(static, 1) num_threads(nt)
        for (long long b = g0; b < g1; b++) {
            std::vector<char>& buf = bufs[(size_t)(b - g0)];
            if (buf.size() < cap) buf.resize(cap);
            char* p = buf.data();
            const long long r1 = (b + 1) * BR < N ? (b + 1) * BR : N;
            for (long long i = b * BR; i < r1; i++) {
                for (int d = 0; d < D; d++) { p = fmt_f6(p, ev[(size_t)i * D + d]); if (d + 1 < D) *p++ = ','; }
                *p++ = '\t';
                for (int k = 0; k < K; k++) { p = fmt_f6(p, c->memberships[(size_t)k * N + i]); if (k + 1 < K) *p++ = ','; }
                *p++ = '\n';
            }
            used[(size_t)(b - g0)] = (size_t)(p - buf.data());
        }
        for (long long b = g0; b < g1; b++)
            if (std::fwrite(bufs[(size_t)(b - g0)].data(), 1, used[(size_t)(b - g0)], f) != used[(size_t)(b - g0)]) { ok = false; break; }
    }
    if (std::fclose(f) != 0) ok = false;
    return ok ? GMM_OK : gmm::fail(GMM_ERR_IO, std::string("write error on '") + path + "'");
}

// Header of a "*.bin" input (readData.cpp:35-40): int32 nevents, int32 ndims.
int gmm_read_bin_header(const char* path, int* ndims, int* nevents) {
    if (!path || !ndims || !nevents) return gmm::fail(GMM_ERR_ARG, "gmm_read_bin_header: bad argument");
    FILE* f = std::fopen(path, "rb");
    if (!f) return gmm::fail(GMM_ERR_IO, std::string("cannot open ") + path);
    int32_t hdr[2];
    const bool ok = std::fread(hdr, sizeof(int32_t), 2, f) == 2 && hdr[0] > 0 && hdr[1] > 0;
    std::fclose(f);
    if (!ok) return gmm::fail(GMM_ERR_IO, "bad .bin header");
    *nevents = hdr[0];
    *ndims = hdr[1];
    return GMM_OK;
}

}  // extern "C"
