// tb_simt.h -- the handful of SIMT primitives the bar-lane kernel uses.
//
// Device build (nvcc, sm_100a): thin wrappers over the warp intrinsics, always with the FULL mask -- the kernel keeps
// its control flow warp-uniform at every exchange point (loops run while ANY env of the warp still needs them, idle
// envs are predicated off), so the 32 lanes of a warp never have to re-converge through a partial barrier.
//
// Host build (-DTB_EMUL, tests/emul only): the same calls land in a fibre-based warp emulator (tests/emul/tb_emul.cpp)
// that runs 32 lanes as cooperative fibres and implements shuffle / vote / barrier as rendezvous points.  The product
// library never builds or links the emulator.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__) && !defined(TB_EMUL)
#define TB_DEV 1
#define TB_FN __device__ __forceinline__
#define TB_NOINL __device__ __noinline__
#define TB_UNROLL1 _Pragma("unroll 1")
#define TB_UNROLL _Pragma("unroll")
namespace tb {
TB_FN int simt_lane() { return (int)(threadIdx.x & 31u); }
TB_FN double shfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
TB_FN float shfl(float v, int src) { return __shfl_sync(0xffffffffu, v, src); }
TB_FN int shfl(int v, int src) { return __shfl_sync(0xffffffffu, v, src); }
TB_FN bool any(bool p) { return __any_sync(0xffffffffu, p) != 0; }
TB_FN unsigned ballot(bool p) { return __ballot_sync(0xffffffffu, p); }
TB_FN void wsync() { __syncwarp(0xffffffffu); }
TB_FN int lowbit(unsigned x) { return __ffs((int)x) - 1; }
}  // namespace tb
#else
#define TB_DEV 0
#define TB_FN static inline
#define TB_NOINL static
#define TB_UNROLL1
#define TB_UNROLL
namespace tb {
// implemented by the emulator runtime
int emu_lane();
uint64_t emu_shfl(uint64_t bits, int src);
unsigned emu_ballot(bool p);
void emu_sync();
TB_FN int simt_lane() { return emu_lane(); }
TB_FN double shfl(double v, int src) { uint64_t b; memcpy(&b, &v, 8); b = emu_shfl(b, src); memcpy(&v, &b, 8); return v; }
TB_FN float shfl(float v, int src) { uint64_t b = 0; memcpy(&b, &v, 4); b = emu_shfl(b, src); memcpy(&v, &b, 4); return v; }
TB_FN int shfl(int v, int src) { uint64_t b = (uint32_t)v; b = emu_shfl(b, src); return (int)(uint32_t)b; }
TB_FN bool any(bool p) { return emu_ballot(p) != 0; }
TB_FN unsigned ballot(bool p) { return emu_ballot(p); }
TB_FN void wsync() { emu_sync(); }
TB_FN int lowbit(unsigned x) { return __builtin_ffs((int)x) - 1; }
}  // namespace tb
#endif
