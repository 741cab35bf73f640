// mpb_prefilter.h — internal interface of the bit-sliced entropy prefilter (mpb_prefilter.cu)
#pragma once
#include <stdint.h>

#include "mpb200.h"
#include "mpb_host.h"

// (sum c, sum c log2 c) of every window's code histogram; the caller has validated the arguments, every window lies
// inside the alignment (win_pos + k <= n_col) and no row is shorter than the alignment.  s0 / s1: host or device.
int mpb_prefilter_bs(mpb_msa* m, int k, int v, const int32_t* win_pos, int32_t nw, double* s0_hd, double* s1_hd);
