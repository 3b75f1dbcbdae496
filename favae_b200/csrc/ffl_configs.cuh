// Kernel configurations of the fused spectrum-loss kernel, shared by the CUDA build
// and the host emulation (tests/emul/ffl_emul.cpp).
#pragma once
#include "ffl_core.cuh"

namespace favae {
//                      N    C  MPC  THREADS
using FflCfg8   = FflCfg<8,   1, 32, 128>;
using FflCfg16  = FflCfg<16,  1, 16, 128>;
using FflCfg32  = FflCfg<32,  1, 4,  256>;
using FflCfg64  = FflCfg<64,  1, 1,  256>;
using FflCfg128 = FflCfg<128, 1, 1,  512>;
using FflCfg256 = FflCfg<256, 2, 1,  512>;   // 2-CTA cluster, 1 CTA per SM (256 threads: 5 % slower)
using FflCfg256c4 = FflCfg<256, 4, 1, 256>;  // 4-CTA cluster, 2 CTAs (two maps) per SM
using FflCfg512 = FflCfg<512, 8, 1,  256>;   // 8-CTA cluster (128 KB of spectrum per CTA), 32 values per thread
}  // namespace favae
