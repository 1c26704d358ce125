// hx_intpeak.cu — measurement helpers: the integer-ALU roofline denominator.
//
// MEASURED_PEAKS.json has no integer figure, so the Myers kernel's roofline
// denominator is measured on the box by these micro-kernels: independent dependent-chains of one
// SASS instruction class per variant, all SMs, enough warps to saturate the issue port.
//   0 LOP3            (alu pipe)          1 IADD3           (alu pipe)
//   2 SHF             (alu pipe)          3 IMAD            (fma pipe)
//   4 LOP3+IMAD 1:1   (both pipes)        5 IMAD.HI         (fma pipe, wide)
//   6 LOP3+IADD3 1:1  (alu pipe)          7 LOP3+IMAD 2:1
//   10..13 the Myers step itself (variants 0..3 of hx_kernels.cu) on register operands:
//            result is in word-steps, not instructions.
#include "../../include/hyperex_b200.h"
#include "../../include/hyperex_b200_bench.h"

#include <cuda_runtime.h>
#include <stdint.h>

namespace {

constexpr int CHAINS = 8;
constexpr int UNROLL = 32;

template <int VAR>
__device__ __forceinline__ void op(uint32_t &x, const uint32_t y, const uint32_t z, const int chain, const uint32_t nb) {
    if (VAR == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(y), "r"(z));
    else if (VAR == 1) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(nb));
    else if (VAR == 2) asm volatile("shf.l.wrap.b32 %0, %0, %1, 5;" : "+r"(x) : "r"(y));
    else if (VAR == 3) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(z));
    else if (VAR == 4) {
        if (chain & 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(z));
        else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(y), "r"(z));
    } else if (VAR == 5) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(z));
    else if (VAR == 6) {
        if (chain & 1) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(nb));
        else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(y), "r"(z));
    } else if (VAR == 9) {
        uint32_t t;
        asm volatile("popc.b32 %0, %1;" : "=r"(t) : "r"(x));
        x = t | y;  // keeps the chain dependent; the OR is a LOP3 (ALU pipe)
    } else if (VAR == 8) {
        uint64_t acc = ((uint64_t)z << 32) | x;
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(x), "r"(y));
        x = (uint32_t)acc ^ (uint32_t)(acc >> 32);
    } else if (VAR == 7) {
        if (chain % 3 == 2) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(z));
        else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(y), "r"(z));
    }
}

template <int VAR>
__global__ void __launch_bounds__(256) hx_intpeak_kernel(uint32_t *out, int iters, uint32_t y, uint32_t z) {
    uint32_t x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x * 2654435761u + c;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) op<VAR>(x[c], y, z, c, x[(c + 3) & (CHAINS - 1)]);
    }
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc ^= x[c];
    if (acc == 0x12345678u) out[0] = acc;  // never true in practice; keeps the chains alive
}

template <int V>
__device__ __forceinline__ void step(const uint32_t eq, uint32_t &pv, uint32_t &mv, int &s, const uint32_t two) {
    const uint32_t xv = eq | mv;
    const uint32_t xh = (((eq & pv) + pv) ^ pv) | eq;
    uint32_t ph = mv | ~(xh | pv);
    uint32_t mh = pv & xh;
    if (V == 3) {
        asm("mad.hi.u32 %0, %1, %2, %0;" : "+r"(s) : "r"(ph), "r"(two));
        asm("mad.hi.s32 %0, %1, %2, %0;" : "+r"(s) : "r"(mh), "r"(two));
        ph <<= 1;
        mh <<= 1;
    } else if (V == 1) {
        asm("mad.hi.u32 %0, %1, 2, %0;" : "+r"(s) : "r"(ph));
        asm("mad.hi.s32 %0, %1, 2, %0;" : "+r"(s) : "r"(mh));
        ph <<= 1;
        mh <<= 1;
    } else if (V == 2) {
        asm("add.cc.u32 %0, %0, %0;\n\taddc.s32 %1, %1, 0;" : "+r"(ph), "+r"(s));
        asm("add.cc.u32 %0, %0, %0;\n\tsubc.s32 %1, %1, 0;" : "+r"(mh), "+r"(s));
    } else {
        s += (int)(ph >> 31) - (int)(mh >> 31);
        ph <<= 1;
        mh <<= 1;
    }
    pv = mh | ~(xv | ph);
    mv = ph & xv;
}

constexpr int NPAT = 15;

// Variant 4: `ph << 1` and `score += ph >> 31` in ONE FMA-pipe instruction: the 64-bit multiply-add
// (0:ph) * one + (sP:ph) = (sP + carry : 2 ph).  Same for mh with its own accumulator sM; a hit is
// sP - sM < 0 with sP initialised to m - k - 1.
__device__ __forceinline__ void step_wide(const uint32_t eq, uint32_t &pv, uint32_t &mv, uint32_t &sP, uint32_t &sM,
                                          const uint32_t one) {
    const uint32_t xv = eq | mv;
    const uint32_t xh = (((eq & pv) + pv) ^ pv) | eq;
    uint32_t ph = mv | ~(xh | pv);
    uint32_t mh = pv & xh;
    uint64_t a, b;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "r"(ph), "r"(sP));
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a) : "r"(ph), "r"(one));
    asm("mov.b64 {%0, %1}, %2;" : "=r"(ph), "=r"(sP) : "l"(a));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "r"(mh), "r"(sM));
    asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(b) : "r"(mh), "r"(one));
    asm("mov.b64 {%0, %1}, %2;" : "=r"(mh), "=r"(sM) : "l"(b));
    pv = mh | ~(xv | ph);
    mv = ph & xv;
}

__global__ void __launch_bounds__(128, 4) hx_steppeak_wide_kernel(uint32_t *out, int iters, uint32_t seed, uint32_t one) {
    uint32_t pv[NPAT], mv[NPAT], sP[NPAT], sM[NPAT];
#pragma unroll
    for (int p = 0; p < NPAT; ++p) {
        pv[p] = ~0u;
        mv[p] = 0;
        sP[p] = 20;
        sM[p] = 0;
    }
    uint32_t e = seed ^ (threadIdx.x * 2654435761u);
    int any = 0;
    for (int it = 0; it < iters; ++it) {
        e = e * 1664525u + 1013904223u;
#pragma unroll
        for (int p = 0; p < NPAT; ++p) {
            step_wide(e ^ (0x9e3779b9u * (p + 1)), pv[p], mv[p], sP[p], sM[p], one);
            any |= (int)(sP[p] - sM[p]);
        }
    }
    uint32_t acc = (uint32_t)any;
#pragma unroll
    for (int p = 0; p < NPAT; ++p) acc ^= pv[p] ^ mv[p];
    if (acc == 0x12345678u) out[0] = acc;
}

// Variant 5: no per-step score at all; dist = popc(pv) - popc(mv) (sum of the vertical deltas) is recomputed
// from the state for the hit test, moving the two score instructions from the ALU pipe to POPC.
__global__ void __launch_bounds__(128, 4) hx_steppeak_popc_kernel(uint32_t *out, int iters, uint32_t seed) {
    uint32_t pv[NPAT], mv[NPAT];
#pragma unroll
    for (int p = 0; p < NPAT; ++p) {
        pv[p] = ~0u;
        mv[p] = 0;
    }
    uint32_t e = seed ^ (threadIdx.x * 2654435761u);
    int any = 0;
    for (int it = 0; it < iters; ++it) {
        e = e * 1664525u + 1013904223u;
#pragma unroll
        for (int p = 0; p < NPAT; ++p) {
            const uint32_t eq = e ^ (0x9e3779b9u * (p + 1));
            const uint32_t xv = eq | mv[p];
            const uint32_t xh = (((eq & pv[p]) + pv[p]) ^ pv[p]) | eq;
            uint32_t ph = mv[p] | ~(xh | pv[p]);
            uint32_t mh = pv[p] & xh;
            ph <<= 1;
            mh <<= 1;
            pv[p] = mh | ~(xv | ph);
            mv[p] = ph & xv;
            any |= __popc(pv[p]) - __popc(mv[p]) - (13 + p);
        }
    }
    uint32_t acc = (uint32_t)any;
#pragma unroll
    for (int p = 0; p < NPAT; ++p) acc ^= pv[p] ^ mv[p];
    if (acc == 0x12345678u) out[0] = acc;
}

template <int V>
__global__ void __launch_bounds__(128, 4) hx_steppeak_kernel(uint32_t *out, int iters, uint32_t seed, uint32_t two) {
    uint32_t pv[NPAT], mv[NPAT];
    int s[NPAT];
#pragma unroll
    for (int p = 0; p < NPAT; ++p) {
        pv[p] = ~0u;
        mv[p] = 0;
        s[p] = 20;
    }
    uint32_t e = seed ^ (threadIdx.x * 2654435761u);
    int any = 0;
    for (int it = 0; it < iters; ++it) {
        e = e * 1664525u + 1013904223u;
#pragma unroll
        for (int p = 0; p < NPAT; ++p) {
            step<V>(e ^ (0x9e3779b9u * (p + 1)), pv[p], mv[p], s[p], two);
            any |= s[p];
        }
    }
    uint32_t acc = (uint32_t)any;
#pragma unroll
    for (int p = 0; p < NPAT; ++p) acc ^= pv[p] ^ mv[p];
    if (acc == 0x12345678u) out[0] = acc;
}

template <int VAR>
cudaError_t run_peak(uint32_t *d, int iters, float *ms) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const int blocks = 148 * 8;
    hx_intpeak_kernel<VAR><<<blocks, 256>>>(d, iters / 8, 0x0f0f1234u, 0x33cc5678u);  // warm-up
    cudaEventRecord(a);
    hx_intpeak_kernel<VAR><<<blocks, 256>>>(d, iters, 0x0f0f1234u, 0x33cc5678u);
    cudaEventRecord(b);
    cudaError_t e = cudaEventSynchronize(b);
    cudaEventElapsedTime(ms, a, b);
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    return e;
}

template <int V>
cudaError_t run_step(uint32_t *d, int iters, float *ms) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const int blocks = 148 * 4 * 4;
    if (V == 4) hx_steppeak_wide_kernel<<<blocks, 128>>>(d, iters / 8, 12345u, 1u);
    else if (V == 5) hx_steppeak_popc_kernel<<<blocks, 128>>>(d, iters / 8, 12345u);
    else hx_steppeak_kernel<V><<<blocks, 128>>>(d, iters / 8, 12345u, 2u);
    cudaEventRecord(a);
    if (V == 4) hx_steppeak_wide_kernel<<<blocks, 128>>>(d, iters, 12345u, 1u);
    else if (V == 5) hx_steppeak_popc_kernel<<<blocks, 128>>>(d, iters, 12345u);
    else hx_steppeak_kernel<V><<<blocks, 128>>>(d, iters, 12345u, 2u);
    cudaEventRecord(b);
    cudaError_t e = cudaEventSynchronize(b);
    cudaEventElapsedTime(ms, a, b);
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    return e;
}

}  // namespace

// Returns in *gops the measured throughput in 1e9 lane-operations per second (variants 0-7:
// integer instructions x 32 lanes; variants 10-12: Myers word-steps).  iters scales the run time.
extern "C" int hx_int_peak(int device, int variant, int iters, double *gops) {
    if (!gops || iters < 8) return HX_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return HX_ERR_CUDA;
    uint32_t *d = nullptr;
    if (cudaMalloc((void **)&d, 256) != cudaSuccess) return HX_ERR_CUDA;
    float ms = 0;
    cudaError_t e = cudaSuccess;
    double ops = 0;
    switch (variant) {
    case 0: e = run_peak<0>(d, iters, &ms); break;
    case 1: e = run_peak<1>(d, iters, &ms); break;
    case 2: e = run_peak<2>(d, iters, &ms); break;
    case 3: e = run_peak<3>(d, iters, &ms); break;
    case 4: e = run_peak<4>(d, iters, &ms); break;
    case 5: e = run_peak<5>(d, iters, &ms); break;
    case 6: e = run_peak<6>(d, iters, &ms); break;
    case 7: e = run_peak<7>(d, iters, &ms); break;
    case 8: e = run_peak<8>(d, iters, &ms); break;
    case 9: e = run_peak<9>(d, iters, &ms); break;
    case 10: e = run_step<0>(d, iters, &ms); break;
    case 11: e = run_step<1>(d, iters, &ms); break;
    case 12: e = run_step<2>(d, iters, &ms); break;
    case 13: e = run_step<3>(d, iters, &ms); break;
    case 14: e = run_step<4>(d, iters, &ms); break;
    case 15: e = run_step<5>(d, iters, &ms); break;
    default: cudaFree(d); return HX_ERR_ARG;
    }
    if (variant < 10) ops = (double)148 * 8 * 256 * (double)iters * UNROLL * CHAINS;
    else ops = (double)148 * 16 * 128 * (double)iters * NPAT;
    cudaFree(d);
    if (e != cudaSuccess || ms <= 0) return HX_ERR_CUDA;
    *gops = ops / (ms * 1e-3) / 1e9;
    return HX_OK;
}
