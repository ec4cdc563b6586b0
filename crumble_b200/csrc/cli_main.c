/* crumble_gpu: the crumble command line (reference main(), snp_score.c:2144) on the GPU path */
#include "crumble_host.h"
int main(int argc, char **argv) { return crumble_main(argc, argv); }
