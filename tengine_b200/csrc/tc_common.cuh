// tc_common.cuh -- device helpers shared by the tcgen05 kernels (gemm_tcgen05.cu, conv_window.cu): UMMA shared-memory and
// instruction descriptors, the SW32 tile addressing of the thread-built A tiles, the per-lane fast epilogue unit.
#pragma once
#include "common.cuh"
#include "ptx.cuh"

namespace tb200 {

static constexpr int BLOCK_M = 128;

// K-major operand tile in shared memory, rows of `swizzle` bytes, 8-row groups `8*swizzle` bytes apart
// (cute/atom/mma_traits_sm100.hpp: canonical layout ((8,n),2):((swizzle/16,SBO),1), LBO = 1, version 1).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, int swizzle)
{
    const uint64_t layout = (swizzle == 128) ? 2ull : (swizzle == 64) ? 4ull : 6ull;
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
    d |= (uint64_t)1 << 16;                              // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(((8 * swizzle) >> 4) & 0x3fff) << 32; // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                              // descriptor version (Blackwell)
    d |= layout << 61;
    return d;
}

// UMMA instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): S32 accumulate, A/B int8 or uint8,
// both K-major, M = 128, N = block_n.
__host__ __device__ inline uint32_t make_idesc_i8(int block_n, bool a_signed, bool b_signed)
{
    uint32_t d = 0;
    d |= 2u << 4;                        // c_format = S32
    d |= (a_signed ? 1u : 0u) << 7;      // a_format
    d |= (b_signed ? 1u : 0u) << 10;     // b_format
    d |= (uint32_t)(block_n >> 3) << 17; // n_dim
    d |= (uint32_t)(BLOCK_M >> 4) << 24; // m_dim
    return d;
}

template <bool FUSE>
__device__ __forceinline__ void stem_unit_fast(const uint32_t (&v)[16], uint32_t par_addr, int oc0, const EpiParams& e, uint32_t (&w)[4])
{
    float gw[4];
#pragma unroll
    for (int h = 0; h < 2; h++)
    {
        float4 p[4];
#pragma unroll
        for (int k = 0; k < 4; k++) p[k] = lds_f4(par_addr + h * 64 + k * 16);
        const int32_t a8[8] = {(int32_t)v[h * 8], (int32_t)v[h * 8 + 1], (int32_t)v[h * 8 + 2], (int32_t)v[h * 8 + 3],
                               (int32_t)v[h * 8 + 4], (int32_t)v[h * 8 + 5], (int32_t)v[h * 8 + 6], (int32_t)v[h * 8 + 7]};
        requant_fast8_i8<FUSE>(a8, p, e, w[2 * h], w[2 * h + 1], gw[2 * h], gw[2 * h + 1]);
    }
    if (e.q_byte_add)
    {
#pragma unroll
        for (int j = 0; j < 4; j++) w[j] = requant_byte_fix(w[j], e);
    }
    if (fmaxf(fmaxf(gw[0], gw[1]), fmaxf(gw[2], gw[3])) > 0.5f - TB200_TIE_EPS)
    {
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (gw[j] > 0.5f - TB200_TIE_EPS)
                w[j] = requant_fix_word<FUSE>(w[j], (int32_t)v[j * 4], (int32_t)v[j * 4 + 1], (int32_t)v[j * 4 + 2], (int32_t)v[j * 4 + 3], oc0 + j * 4, e);
    }
}

// row r, 16-byte chunk c16 of a SW32 K-major tile (8-row groups of 256 bytes, chunk index ^= bit 2 of the row)
__device__ __forceinline__ uint32_t sw32_offset(int r, int c16) { return (uint32_t)((r >> 3) * 256 + (r & 7) * 32 + ((c16 ^ ((r >> 2) & 1)) << 4)); }


} // namespace tb200
