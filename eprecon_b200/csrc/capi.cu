#include "common.cuh"
extern "C" int ep_version(void) { return 100; }
