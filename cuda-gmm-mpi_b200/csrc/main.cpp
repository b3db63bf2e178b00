// Entry point of the drop-in executable (reference: gaussianMPI, Makefile:37).
#include "../../include/gmm.h"
int main(int argc, char** argv) { return gmm_main(argc, argv); }
