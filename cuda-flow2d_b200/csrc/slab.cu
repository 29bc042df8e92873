// slab.cu -- halo exchange of the row-slab decomposition (one large frame on several GPUs): direct stores into the
// neighbour GPU's memory over NVLink plus a flag, no host involvement, no NCCL call on the data path.  sm_100a.
//
// No counterpart upstream (the reference is single-GPU: cuDeviceGet(.., 0), cuda_operation_solve_2d.cpp:168);
// SURVEY.md 8(e), BASELINE.json configs[4].
//
// Every rank owns a MAILBOX (one cudaMalloc, mapped into its two neighbours by peer access or CUDA IPC):
//     flags[2]                              epoch of the last complete message from the rank above / below
//     recv[from above | from below][epoch & 1][field 0 | 1][rows x pitch]
// A message = some rows of two fields (du, dv or u, v).  slab_push_kernel copies the sender's rows into the
// receiver's mailbox (st.global to the peer mapping), fences at system scope, and the last CTA to finish publishes
// the epoch.  In stream order on the receiver, slab_wait_kernel spins on the flag (ld.acquire.sys) and slab_unpack_kernel
// copies the rows from the mailbox into the receiver's container.  Epochs are consumed in order and messages alternate between
// two buffers; a sender can only be two epochs ahead after having received the answer to the previous one, so a
// buffer is never overwritten while it is being read.
#include "kernels.h"

namespace flow2d {

constexpr unsigned long long kSlabTimeoutNs = 20ull * 1000 * 1000 * 1000;  // 20 s

__global__ void __launch_bounds__(256)
slab_push_kernel(SlabPush p) {
  // blockIdx.z = direction (0: to the rank above, 1: to the rank below)
  const SlabPushDir& d = p.dir[blockIdx.z];
  if (d.rows > 0) {
    const int x = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
    for (int r = blockIdx.y; r < d.rows; r += gridDim.y) {
      if (x < p.w) {
        const size_t src = (size_t)(d.row0 + r) * p.pitch + x, dst = (size_t)r * p.pitch + x;
        // whole float4s: containers are pitch-padded (pitch % 4 == 0), the padding is never read as data
        *reinterpret_cast<float4*>(d.dst[0] + dst) = *reinterpret_cast<const float4*>(p.field[0] + src);
        *reinterpret_cast<float4*>(d.dst[1] + dst) = *reinterpret_cast<const float4*>(p.field[1] + src);
      }
    }
  }
  // the last CTA of the launch publishes the epochs (release at system scope: the stores above are visible to
  // the neighbour before the flag is)
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned total = gridDim.x * gridDim.y * gridDim.z;
    if (atomicAdd(p.counter, 1u) == total - 1) {
      *p.counter = 0;
      __threadfence_system();
      for (int k = 0; k < 2; k++)
        if (p.dir[k].rows > 0)
          asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p.dir[k].flag), "l"(p.dir[k].epoch) : "memory");
    }
  }
}

// One thread per direction waits for the neighbour's epoch; the copy kernel follows in stream order.  (A single tiny
// CTA spins: with the wait inside the copy kernel its whole grid would sit on the SMs, and several logical ranks on
// one GPU -- the tests -- could starve the very kernels they are waiting for.)
__global__ void __launch_bounds__(32)
slab_wait_kernel(SlabUnpack p) {
  if (threadIdx.x >= 2) return;
  const SlabUnpackDir& d = p.dir[threadIdx.x];
  if (d.rows <= 0) return;
  // A neighbour that never sends (a rank that failed, a device-wide synchronisation on the host while another rank's
  // kernels still have to be enqueued) must not hang the GPU: give up after kSlabTimeoutNs and report it.
  unsigned long long seen, t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while (true) {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(d.flag) : "memory");
    if (seen >= d.epoch) break;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > kSlabTimeoutNs) {
      atomicExch(p.error, 1u);
      break;
    }
    __nanosleep(100);
  }
}

__global__ void __launch_bounds__(256)
slab_unpack_kernel(SlabUnpack p) {
  const SlabUnpackDir& d = p.dir[blockIdx.z];
  if (d.rows <= 0) return;
  const int x = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
  if (x >= p.w) return;
  for (int r = blockIdx.y; r < d.rows; r += gridDim.y) {
    const size_t src = (size_t)r * p.pitch + x, dst = (size_t)(d.row0 + r) * p.pitch + x;
    // ld.volatile: the mailbox was written by another GPU; nothing of it may come from a stale L1 line
    float4 a, b;
    asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(d.src[0] + src));
    asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(d.src[1] + src));
    *reinterpret_cast<float4*>(p.field[0] + dst) = a;
    *reinterpret_cast<float4*>(p.field[1] + dst) = b;
  }
}

void launch_slab_push(cudaStream_t st, const SlabPush& p) {
  int rows = p.dir[0].rows > p.dir[1].rows ? p.dir[0].rows : p.dir[1].rows;
  if (rows <= 0) return;
  dim3 block(256), grid((p.w + 1023) / 1024, rows < 64 ? rows : 64, 2);
  slab_push_kernel<<<grid, block, 0, st>>>(p);
}

void launch_slab_unpack(cudaStream_t st, const SlabUnpack& p) {
  int rows = p.dir[0].rows > p.dir[1].rows ? p.dir[0].rows : p.dir[1].rows;
  if (rows <= 0) return;
  dim3 block(256), grid((p.w + 1023) / 1024, rows < 64 ? rows : 64, 2);
  slab_wait_kernel<<<1, 32, 0, st>>>(p);
  slab_unpack_kernel<<<grid, block, 0, st>>>(p);
}

void preload_slab_kernels() {
  cudaFuncAttributes a;
  cudaFuncGetAttributes(&a, slab_push_kernel);
  cudaFuncGetAttributes(&a, slab_wait_kernel);
  cudaFuncGetAttributes(&a, slab_unpack_kernel);
}

}  // namespace flow2d
