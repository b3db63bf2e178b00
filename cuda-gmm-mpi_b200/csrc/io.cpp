// io.cpp — input readers and output writers with the reference's file
// formats (readData.cpp:25-129; gaussian.cu:998-1061, 1180-1201).  One-shot
// I/O: kept format-compatible, not accelerated.
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "host_math.h"

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

// .results (gaussian.cu:1042-1059): "x1,...,xD<TAB>g1,...,gK\n", all %f;
// memberships are cluster-major [K][N].
int gmm_write_results(const char* path, const float* ev, long long N, int D, const clusters_t* c, int K) {
    FILE* f = std::fopen(path, "w");
    if (!f) return gmm::fail(GMM_ERR_IO, std::string("Unable to open file '") + path + "' for writing.");
    std::vector<char> buf(1 << 20);
    std::setvbuf(f, buf.data(), _IOFBF, buf.size());
    for (long long i = 0; i < N; i++) {
        for (int d = 0; d < D; d++) std::fprintf(f, d + 1 < D ? "%f," : "%f", ev[(size_t)i * D + d]);
        std::fputc('\t', f);
        for (int k = 0; k < K; k++) std::fprintf(f, k + 1 < K ? "%f," : "%f", c->memberships[(size_t)k * N + i]);
        std::fputc('\n', f);
    }
    std::fclose(f);
    return GMM_OK;
}

}  // extern "C"
