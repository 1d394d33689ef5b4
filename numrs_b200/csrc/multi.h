// multi.h -- one host-slice call spread over several GPUs of the box, inside ONE process (no torch, no IPC, no NCCL).
//
// The reference's entry points take whole host arrays (`rlft3(&mut Array3, ..)` Real_FT3.rs:8, the in-memory `Fourn`
// call Real_FT3.rs:35, `fft_batch` FFT_1.rs:185, `convlv_batch` Convolve.rs:241, `correl_batch` Correlation.rs:273), so a
// drop-in can only use more than one GPU if the library itself scatters the work.  With the option `num_devices` > 1
// (or 0 = every visible device) the C ABI does that:
//   * 3-D transforms: the host volume is scattered as nn2-slabs (forward) / nn1-slabs (inverse) by strided H2D copies on
//     one stream per device -- the scatter realises the slab input for free --, every device runs the slab
//     programs of plan.cpp with the exchange fused into the FFT epilogue (stores into the peers' receive buffers over
//     NVLink, cudaDeviceEnablePeerAccess), cross-device events order stage 1 behind every device's stage 0, and the
//     result slabs are gathered by D2H copies.  All the PCIe links of the box work at once.
//   * batches: contiguous batch ranges per device, no communication.
// A persistent worker thread per device issues that device's copies and launches (pageable host memory makes
// cudaMemcpyAsync synchronous, so one issuing thread would serialise the links).
#pragma once
#include <functional>
#include <string>

#include "plan.h"

namespace nrb {

// devices one host-slice call may use: `num_devices` clamped to the visible devices and rounded down to a power of two
// (<= 8); 1 when the option is 1 (default)
int multi_device_count();

// runs fn(g) for g = 0 .. G-1 concurrently, worker g with device g current; returns the first failing rc and copies that
// worker's error message into the caller's nrb_last_error()
int multi_run(int G, const std::function<int(int)> &fn);

// whole-array 3-D transform over G devices (data / speq are host pointers; speq == nullptr and real == false: complex
// fourn).  Returns NRB_ERR_UNSUPPORTED when the shape or the box cannot take the slab path (the caller then uses one device).
int multi_transform3d(bool real, double *data, double *speq, size_t nn1, size_t nn2, size_t nn3, int isign, int G);

// contiguous batch ranges over the devices: fn(first, count) runs on a worker whose device is current; ranges are
// ceil(total / G) long.  Serialised with the other multi-device calls.
int multi_shard_batch(size_t total, int G, const std::function<int(size_t, size_t)> &fn);

// calls that really took the multi-device path since the library was loaded (tests: a silent single-device fallback
// would otherwise look like success): which = 0 3-D transforms, 1 sharded batches
long multi_calls(int which);

// drops the cached multi-device plans and their device buffers (nrb_shutdown, nrb_set_option)
void multi_release();

} // namespace nrb
