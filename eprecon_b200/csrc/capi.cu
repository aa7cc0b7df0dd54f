#include "common.cuh"

extern "C" {

int ep_version(void) { return 100; }

// How host threads wait in cudaStreamSynchronize on the CURRENT device: 0 = driver default (spins when cores are free), 1 = spin,
// 2 = yield, 4 = block on an OS primitive.  One fragment takes ~25 stream drains (segment counts the host needs to size the next
// allocation); with S fragments in flight per GPU and one process per GPU, S x N spinning threads on a box with fewer cores
// slow every rank down (measured: 8 ranks x 8 streams on 32 cores, 0.67 scaling efficiency) -- blocking waits give the cores
// back.  Applies to the device's primary context, i.e. to torch's syncs too.
int ep_set_sync_mode(int mode) {
  unsigned flags;
  switch (mode) {
    case 0: flags = cudaDeviceScheduleAuto; break;
    case 1: flags = cudaDeviceScheduleSpin; break;
    case 2: flags = cudaDeviceScheduleYield; break;
    case 4: flags = cudaDeviceScheduleBlockingSync; break;
    default: return EP_ERR_ARG;
  }
  return cudaSetDeviceFlags(flags) == cudaSuccess ? EP_OK : EP_ERR_CUDA;
}

}  // extern "C"
