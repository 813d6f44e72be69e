/*
 * drr_oracle.c -- CPU oracle for the DRR projector hot path (Siddon + trilinear, forward and backward).
 * TEST INFRASTRUCTURE ONLY: a restatement of /root/reference/diffdrr/renderers.py in plain C, in fp32
 * (operation-for-operation like the reference's tensor algebra) and fp64 (the gradient ground truth).
 * Parity pin: every function is checked against outputs of the unmodified reference recorded in
 * the tests/golden fixtures (tests/test_oracle.py).  See drr_oracle_impl.h for the per-function citations.
 *
 * Build: make -C oracle   (gcc -O2 -fopenmp -ffp-contract=off: no FMA contraction, so the fp32 path
 * rounds like ATen's separate mul/add kernels).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define REAL float
#define SUF f32
#include "drr_oracle_impl.h"
#undef REAL
#undef SUF

#define REAL double
#define SUF f64
#include "drr_oracle_impl.h"
#undef REAL
#undef SUF

#ifdef _OPENMP
#include <omp.h>
int oracle_max_threads(void) { return omp_get_max_threads(); }
void oracle_set_threads(int n) { omp_set_num_threads(n); }
#else
int oracle_max_threads(void) { return 1; }
void oracle_set_threads(int n) { (void)n; }
#endif
