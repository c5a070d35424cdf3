// Packed MAF weight image shared by host code and kernels.
//
// One image holds the K masked-MADE conditioners of the flow (reference: K x pyro ConditionalAutoRegressiveNN,
// rbi/rbi/models/transforms.py:92-98) in "order space": theta columns of layer 0 and the rows of the output
// layer are permuted so that position j is the variable that the autoregressive order visits j-th
// (perm[j] = t{k}.nn.permutation[j]).  Pyro's masks then become plain prefixes:
//   layer-0 unit u       sees theta positions  j  < deg0[u]
//   layer-l unit v       sees previous units   u  < pre[l][v]
//   output of position j sees last-layer units v  < preo[j]
// (hidden degrees are non-decreasing in the unit index because they come from round(linspace(1, D, H)) - 1).
// Masked weights are stored as exact zeros, rows are padded to a multiple of 4 floats for 128-bit loads.
#pragma once
#include <stdint.h>

#define RABI_MAXL 4  // hidden layers
#define RABI_MAXK 8  // transforms

struct TImg {
    // float offsets into MafImg::f
    int W0c, ldc;  // [H0p+4][ldc]   context part of layer 0 (bias b0 is folded into the projection)
    int W0t, ldt;  // [H0p+4][ldt]   theta part of layer 0, order-space columns
    int b0;        // [H0p+4]
    int W[RABI_MAXL], ld[RABI_MAXL], b[RABI_MAXL];  // hidden layer l = 1..L-1: [Hlp+4][ld[l]]
    int Wo, ldo, bo;                                // [2D (+pad)][ldo]: row j = mean, row D+j = log-scale of position j
    int WoT, boF;   // unit-major copy of the output layer for k_fast: WoT[v][0..ldt) = mean weights of the positions,
                    // WoT[v][ldt..2ldt) = log-scale weights; boF[0..ldt) / boF[ldt..2ldt) the biases
    int fast_end;   // end (floats) of the contiguous block [W0t, fast_end) staged by k_fast
    int meta_fast_end;  // end (ints) of the contiguous index block [perm, meta_fast_end) staged by k_fast
    // int offsets into MafImg::m
    int perm;             // [D]   variable index of position j on the base-noise side (input of T_k)
    int permy;            // [D]   index of position j on the theta side (output of T_k, after an optional
                          //       Permute layer: permy[j] = inv_permute_k[perm[j]]; == perm without shuffle)
    int deg0;             // [H0]  prefix length over theta positions (== MADE degree of the unit)
    int pre[RABI_MAXL];   // l = 1..L-1: [H_l] prefix length over layer l-1
    int preo;             // [D]
    int grp[RABI_MAXL];   // l = 0..L-1: [D+1] units of degree j are [grp[j], grp[j+1])
    int first[RABI_MAXL]; // l = 1..L-1: [H_{l-1} (+pad)] first consumer row v in layer l of input unit u
    int firsto;           // [H_{L-1} (+pad)] first output position j consuming last-layer unit v
    int first0;           // [D] first layer-0 unit consuming theta position j
};

struct MafImg {
    int K, D, C, L;
    int H[RABI_MAXL];
    int Hmax;
    float lo, hi;     // log-scale clamp
    int W0cT, ldKH;   // [C][ldKH]: transposed context weights of all transforms, column k*H0p + u
    int H0p;
    TImg t[RABI_MAXK];
    const float* f;
    const int* m;
};
