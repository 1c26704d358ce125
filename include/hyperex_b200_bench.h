/*
 * hyperex_b200_bench.h — measurement-only exports of libhyperex_b200.so (not part of the product interface in
 * hyperex_b200.h; nothing on the reference's path corresponds to them).  Used by bench.py.
 */
#ifndef HYPEREX_B200_BENCH_H
#define HYPEREX_B200_BENCH_H
#ifdef __cplusplus
extern "C" {
#endif

/* hx_intpeak.cu: measured throughput, in 1e9 lane-operations per second, of one integer instruction class on
 * `device` (variant 0 LOP3, 1 IADD3, 2 SHF, 3 IMAD, 4 LOP3+IMAD 1:1, 5 IMAD.HI, 6 LOP3+IADD3, 7 LOP3+IMAD 2:1) or
 * of the bare Myers step on register operands (10-12: step variants 0-2, in word-steps).  This is the integer-ALU
 * roofline denominator of bench.py. */
int hx_int_peak(int device, int variant, int iters, double *gops);

#ifdef __cplusplus
}
#endif
#endif /* HYPEREX_B200_BENCH_H */
