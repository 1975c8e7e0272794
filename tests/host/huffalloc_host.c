/* Host build of compressjs_b200/csrc/huffalloc.cuh for the CPU test-suite (no GPU needed). */
#include "../../compressjs_b200/csrc/huffalloc.cuh"
__attribute__((visibility("default"))) void host_ha_allocate(int* a, int len, int maxLen) { ha_allocate(a, len, maxLen); }
