// common.cuh -- device-side numerics shared by every kernel of the B200 backend.
//
// The requantising epilogue reproduces the reference CPU device's float arithmetic operation for
// operation (no FMA contraction, IEEE division, C round() = half away from zero), so that int8 results are
// bit-identical to the reference and uint8 results differ only by the reference's own fp32 summation noise.
// Reference recipes (paths relative to source/device/cpu/op/):
//   HCL int8  : conv/x86/conv_kernel_x86.c:1827-1889, conv/x86/conv_dw_hcl_x86.c:198-262
//   REF int8  : conv/conv_kernel_ref_int8.c:136-171
//   HCL uint8 : conv/x86/conv_kernel_x86.c:1729-1791
//   REF uint8 : conv/conv_kernel_ref_uint8.c:157-186
//   FC int8   : fc/fc_ref.c:225,252      FC uint8 : fc/fc_ref.c:146-166
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/tengine_b200.h"

namespace tb200 {

// Scalar epilogue parameters, passed by value to kernels (per-channel arrays stay in global memory).
struct EpiParams
{
    const int32_t* bias;  // [OCp]
    const float* w_scale; // [OCp]
    float in_scale, out_scale;
    float in_w_scale;     // uint8: s_in * s_w (per tensor), computed once on the host in fp32
    int32_t in_zero, w_zero, out_zero;
    int32_t activation;
    int32_t recipe;
    int32_t is_uint8;
    int32_t fc_rounding;
    int32_t has_bias;
    // ---- fast path (see requant_fast) ----
    const float2* fast_par; // [OCp] int8: (M[oc] ~ s_in*s_w[oc]/s_out [FC: exactly the reference's requant scale], bias bits);
                            //       uint8: (the bias term in real units, exactly as the reference rounds it, 0)
    float fast_lo, fast_hi;   // clamp of t = f/s_out: activation and the +-127 / 0..255 saturation folded together
    float fast_flo, fast_fhi; // uint8: activation clamp in real units
    float fast_r;             // uint8: fl(1/s_out)
    int32_t fast_ok;          // 0: scales are degenerate, always take the exact path
    int32_t fuse_bias;        // int8 conv: fast_par[oc].y holds fl(bias*M) (a float) and t = fma((float)acc, M, .y); see requant_fast_bits
};

// C round(): half away from zero, exact for every float (CUDA's roundf is the exact, slow-path version).
__device__ __forceinline__ float round_half_away(float x) { return roundf(x); }

__device__ __forceinline__ float act_hcl(float f, int activation)
{
    if (activation == 0) f = (f < 0.f) ? 0.f : f;
    if (activation > 0)
    {
        f = (f < 0.f) ? 0.f : f;
        f = (f > 6.f) ? 6.f : f;
    }
    return f;
}

__device__ __forceinline__ float act_ref(float t, int activation)
{
    if (activation >= 0)
    {
        if (t < 0.f && activation != 1) t = 0.f;
        if (t > 1.f && activation == 1) t = 1.f;
        if (t > 6.f && activation == 6) t = 6.f;
        if (t < -1.f && activation == 1) t = -1.f;
    }
    return t;
}

// acc = exact integer dot product (int8: sum x*w; uint8: sum (x-zx)(w-zw) over in-bounds taps), WITHOUT bias.
// Returns the output byte (int8 bit pattern or uint8).
static __device__ __noinline__ int requant(int32_t acc, int oc, const EpiParams& e)
{
    const int32_t b = e.has_bias ? __ldg(e.bias + oc) : 0;
    if (!e.is_uint8)
    {
        const float s_w = __ldg(e.w_scale + oc);
        int q;
        if (e.fc_rounding)
        {
            // fc_ref.c:225 requant = (s_in*s_w)/s_out ; :252 roundf(acc_with_bias * requant)
            const float rq = __fdiv_rn(__fmul_rn(e.in_scale, s_w), e.out_scale);
            q = (int)roundf(__fmul_rn((float)(acc + b), rq));
        }
        else if (e.recipe == TB200_RECIPE_HCL)
        {
            float f = __fmul_rn(__fmul_rn((float)(acc + b), e.in_scale), s_w);
            f = act_hcl(f, e.activation);
            q = (int)round_half_away(__fdiv_rn(f, e.out_scale));
        }
        else
        {
            float f = __fmul_rn((float)(acc + b), __fmul_rn(e.in_scale, s_w));
            f = act_ref(f, e.activation);
            q = (int)round_half_away(__fdiv_rn(f, e.out_scale));
        }
        q = q > 127 ? 127 : q;
        q = q < -127 ? -127 : q;
        return q & 0xff;
    }
    else
    {
        float f = __fmul_rn((float)acc, e.in_w_scale);
        int q;
        if (e.fc_rounding)
        {
            // fc_ref.c:146,162: data = bias*bias_scale + sum ; roundf(data / s_out) + zp  (bias_scale == s_in*s_w)
            if (e.has_bias) f = __fadd_rn(f, __fmul_rn((float)b, e.in_w_scale));
            q = (int)roundf(__fdiv_rn(f, e.out_scale)) + e.out_zero;
        }
        else
        {
            if (e.has_bias)
            {
                // conv_kernel_x86.c:1723,1740: bias * (s_in*s_w) ; conv_kernel_ref_uint8.c:94: (bias * s_in) * s_w
                const float bt = (e.recipe == TB200_RECIPE_HCL)
                                     ? __fmul_rn((float)b, e.in_w_scale)
                                     : __fmul_rn(__fmul_rn((float)b, e.in_scale), __ldg(e.w_scale + oc));
                f = __fadd_rn(f, bt);
            }
            f = (e.recipe == TB200_RECIPE_HCL) ? act_hcl(f, e.activation) : act_ref(f, e.activation);
            q = (int)round_half_away(__fdiv_rn(f, e.out_scale)) + e.out_zero;
        }
        q = q > 255 ? 255 : q;
        q = q < 0 ? 0 : q;
        return q;
    }
}

// Fast requantisation with an exactness guarantee.
//
// The reference rounds t_ref = fl(fl(fl(x*s_in)*s_w)/s_out) (three roundings) half away from zero.  The fast path forms
// t = fl(x*M) with M = fl(s_in*s_w/s_out computed in double): |t - t_ref| <= 5*2^-24*|t| < 4e-5 inside the clamp range
// [-127,127].  Rounding t to nearest (magic-number add) therefore gives the reference's integer unless t lies within
// kTieEps of a half-integer; exactly those elements (about 2.4e-4 of them) are recomputed with the literal reference
// arithmetic (requant()).  Clamps commute with the monotone division/rounding, so activation and saturation are applied
// to t directly.  For FC the reference itself computes roundf(x*rq) and M = rq exactly, so t == t_ref.
// Cost: ~10 instructions per element instead of ~60 (IEEE division + roundf).
#define TB200_MAGIC 12582912.0f // 1.5 * 2^23
#define TB200_TIE_EPS 1.220703125e-4f // 2^-13

// One element: returns the bit pattern of r = t + MAGIC (low byte = the rounded, clamped integer) and ORs `bit` into
// `bad` when t is inside the tie guard band.  ~10 instructions; U8 is a compile-time switch so no per-element branch.
// FUSE (int8 conv layers whose |bias*M| <= 100 output LSBs, decided at prerun): the integer bias add is folded into
// an FMA, t = fma((float)acc, M, fl(bias*M)), one instruction less.  Extra error <= 2^-23*(|t| + |bias*M|) < 4.1e-5,
// total |t - t_ref| < 6.4e-5 -- still inside the 2^-13 guard band, so exactness is preserved by the same argument.
template <bool U8, bool FUSE = false>
__device__ __forceinline__ uint32_t requant_fast_bits(int32_t acc, const EpiParams& e, float m, int32_t b, uint32_t& bad, uint32_t bit)
{
    float t;
    if (!U8 && FUSE)
        t = __fmaf_rn((float)acc, m, __int_as_float(b));
    else if (!U8)
        t = __fmul_rn((float)(acc + b), m);
    else
    {
        // f is bit-identical to the reference's (same operations); only the division is replaced by * fl(1/s_out)
        float f = __fadd_rn(__fmul_rn((float)acc, e.in_w_scale), m);
        f = fminf(fmaxf(f, e.fast_flo), e.fast_fhi);
        t = __fmul_rn(f, e.fast_r);
    }
    t = fminf(fmaxf(t, e.fast_lo), e.fast_hi);
    const float r = __fadd_rn(t, TB200_MAGIC);
    const float d = __fsub_rn(t, __fsub_rn(r, TB200_MAGIC));
    // bad |= bit when |d| > 0.5 - eps : one FSETP (|d| folds into the operand modifier) + one predicated LOP3
    asm("{\n\t.reg .pred p;\n\tsetp.gt.f32 p, %1, %2;\n\t@p or.b32 %0, %0, %3;\n\t}" : "+r"(bad) : "f"(fabsf(d)), "f"(0.5f - TB200_TIE_EPS), "r"(bit));
    return (uint32_t)__float_as_int(r) + (U8 ? (uint32_t)e.out_zero : 0u);
}

// Four consecutive channels -> one packed 32-bit word (3 PRMT).  `bad` receives bits (bit0 << j) for guarded elements.
template <bool U8, bool FUSE = false>
__device__ __forceinline__ uint32_t requant_fast4(const int32_t (&acc)[4], const EpiParams& e, const float (&m)[4], const int32_t (&b)[4],
                                                  uint32_t& bad, uint32_t bit0)
{
    const uint32_t r0 = requant_fast_bits<U8, FUSE>(acc[0], e, m[0], b[0], bad, bit0);
    const uint32_t r1 = requant_fast_bits<U8, FUSE>(acc[1], e, m[1], b[1], bad, bit0 << 1);
    const uint32_t r2 = requant_fast_bits<U8, FUSE>(acc[2], e, m[2], b[2], bad, bit0 << 2);
    const uint32_t r3 = requant_fast_bits<U8, FUSE>(acc[3], e, m[3], b[3], bad, bit0 << 3);
    return __byte_perm(__byte_perm(r0, r1, 0x0040), __byte_perm(r2, r3, 0x0040), 0x5410);
}

// Replace byte j of `word` by the exact result (rare path: ~2.4e-4 of the elements).
__device__ __forceinline__ uint32_t requant_fix_byte(uint32_t word, int j, int32_t acc, int oc, const EpiParams& e)
{
    const uint32_t q = (uint32_t)requant(acc, oc, e) & 0xffu;
    return (word & ~(0xffu << (8 * j))) | (q << (8 * j));
}

// Generic entry for kernels that handle one word (4 channels oc0..oc0+3) at a time with constants in global memory.
// Pad channels (>= oc_limit) have m = 0, b = 0 in fast_par and therefore produce 0.
template <bool U8>
__device__ __forceinline__ uint32_t requant_word(const int32_t (&acc)[4], int oc0, int oc_limit, const EpiParams& e)
{
    if (!e.fast_ok)
    {
        uint32_t w = 0;
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (oc0 + j < oc_limit) w |= ((uint32_t)requant(acc[j], oc0 + j, e) & 0xffu) << (8 * j);
        return w;
    }
    float m[4];
    int32_t b[4];
#pragma unroll
    for (int j = 0; j < 4; j++)
    {
        const float2 p = __ldg(e.fast_par + oc0 + j);
        m[j] = p.x, b[j] = __float_as_int(p.y);
    }
    uint32_t bad = 0;
    uint32_t w = (!U8 && e.fuse_bias) ? requant_fast4<U8, true>(acc, e, m, b, bad, 1u) : requant_fast4<U8, false>(acc, e, m, b, bad, 1u);
    if (U8)
    {
        // pad lanes of uint8 tensors must hold 0 (not the zero point): mask them
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (oc0 + j >= oc_limit) w &= ~(0xffu << (8 * j)), bad &= ~(1u << j);
    }
    if (bad)
    {
#pragma unroll
        for (int j = 0; j < 4; j++)
            if ((bad >> j) & 1u) w = requant_fix_byte(w, j, acc[j], oc0 + j, e);
    }
    return w;
}

__device__ __forceinline__ int dp4a_s8(int a, int b, int c) { return __dp4a(a, b, c); }
__device__ __forceinline__ unsigned dp4a_u8(unsigned a, unsigned b, unsigned c) { return __dp4a(a, b, c); }

static inline int cpad(int c) { return (c + 15) & ~15; }

} // namespace tb200
