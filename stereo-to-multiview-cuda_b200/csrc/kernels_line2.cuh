// kernels_line2.cuh — the cost-volume kernel, second form: one PERSISTENT CTA per SM, a three-stage ring of
// line-segment tiles in shared memory, a producer warp that keeps the ring full and consumer warps that never
// meet at a CTA-wide barrier.
//
// Why (profiles/r1g_line_kernels_ncu_full.txt): k_line alternates, per CTA, between waiting
// for its tile and summing it; three such CTAs per SM leave the issue slots a third idle in the horizontal
// passes and hold the vertical passes at 85 % of the copy bandwidth, and its four-output blocks read every tile
// position from shared memory once per four outputs, which makes the shared-memory pipe as busy as the FP32
// pipe.  Here:
//   * a tile (S outputs + 2*HP halo positions of one line, 512 bytes per position = 128 disparities) is
//     fetched by the PRODUCER WARP with ONE tensor copy (cp.async.bulk.tensor over a 4-D map of the volume
//     [view][row][column][disparity] -> UTMALDG; a column segment is a box of P rows x 1 column x 128 floats,
//     out-of-image positions are zero-filled by the copy engine), completion counted in bytes on an mbarrier,
//     two tiles ahead of the consumers.  Lines are handed out in runs of consecutive segments through one
//     atomic counter, so the CTAs stay balanced whatever the arm statistics of their image region;
//   * the producer warp also turns the arms of the tile's outputs into WINDOW MASKS: for every block of B
//     consecutive outputs and every tile position the block can reach, a bit per output "this position lies
//     inside that output's window" (lanes split the blocks and positions between them);
//   * a CONSUMER WARP owns one block of B outputs.  It walks the union of the block's windows once: one
//     LDS.128 per tile position, then one predicated FADD2 pair per output, the predicates set by a single R2P
//     from the position's mask.  Every accumulator still receives exactly the reference's additions, ascending
//     from 0.0f (d_ca_cross_sum.cu:282-290); a position outside a window is predicated off, never added as zero;
//   * pass 1 (LM_CI_H) computes its tiles instead of loading them: the producer stages the operand words with
//     cp.async, every consumer warp evaluates its share of tile n+1 (packed f32x2 arithmetic: two evaluations per
//     issue slot for the exponential's argument chain) BEFORE it sums tile n, and the 2*HP positions two
//     consecutive segments share are copied inside shared memory, not evaluated twice.  "Tile n+1 complete" is
//     an mbarrier the warps arrive on after their share and wait on a whole summation later.
// One warp = one pixel's 128 disparities (lane q holds disparities 4q..4q+3), as in k_line: this kernel serves
// plans with LP = 32 (num_disp > 64); k_line keeps the narrower ones.
#pragma once
#include <cuda.h>  // CUtensorMap (the type only; the encoder is resolved at run time)

#include "kernels_line.cuh"

namespace s2mv {

constexpr int kL2MaxConsumers = 16;
constexpr int kL2DescRing = 8;
#ifndef S2MV_L2_UNROLL
#define S2MV_L2_UNROLL 4  // positions per trip of the window walk (2: 3 % slower on the fused vertical kernel)
#endif
constexpr int kL2Unroll = S2MV_L2_UNROLL;
constexpr int kL2MaxBlocks = 32;       // output blocks per tile (S / B), at most
constexpr uint32_t kL2PosBytes = 512;  // one tile position: 32 lanes x float4

struct Line2Args {
    LineArgs a;
    int S;                // outputs per tile, a multiple of B
    int HP;               // halo positions either side of a tile (>= usd; 2*HP a multiple of 4)
    int P;                // S + 2*HP
    int tiles_per_line;   // segments along one line
    int tiles_per_claim;  // consecutive segments handed out together
    int claims_per_line;
    int nlines;           // lines per z slice (rows of a horizontal pass / columns of a vertical one)
    int nz;               // view slots x disparity chunks
    int nclaims;          // nz * nlines * claims_per_line
    int *counter;         // work counter of this launch (zeroed by the host before the launch)
    int use_tmap;         // tiles arrive through the tensor map passed next to these arguments (else 512-byte bulk copies)
    int tmap_rows;        // positions per tensor copy (the map's box): a tile is P / tmap_rows copies
    int row_bias;         // tensor-map row of image row r = r - row_bias (row bands: the volume starts at row v_lo)
};

struct Line2Desc {
    int valid, ln, seg, vslot, chunk, t0, Sact, first;
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "L2_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra L2_DONE;\n"
        "bra L2_WAIT;\n"
        "L2_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
// all cp.async of this thread issued so far -> one arrival on the barrier when they have landed
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
// 4-byte asynchronous copy; src_bytes = 0 writes a zero word
__device__ __forceinline__ void cp_async4(uint32_t dst, const void *src, uint32_t src_bytes)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// bulk asynchronous copy global -> shared (UBLKCP), completion reported in bytes on the barrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// 4-D tiled tensor copy global -> shared (UTMALDG): box fixed by the map, coordinates innermost first
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, int c0, int c1, int c2, int c3, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
        : "memory");
}
__device__ __forceinline__ uint32_t lds16(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

// shared-memory carve-up of one CTA (every section 16-byte aligned); the kernel walks the same list
__host__ __device__ inline size_t line2_align16(size_t b) { return (b + 15) & ~(size_t)15; }
__host__ __device__ inline int line2_frame(int B, int HP) { return (B + 2 * HP + 1) & ~1; }  // mask positions per block
__host__ __device__ inline size_t line2_smem_bytes(int S, int HP, int B, bool ci, int kL2Stages, int LP = 32)
{
    const size_t P = (size_t)S + 2 * HP, NBT = (size_t)(S + B - 1) / B;
    size_t b = 128;                                                                     // alignment slack of the tile base
    b += kL2Stages * P * (size_t)(16 * LP);                                             // tiles
    b += 128;                                                                           // mbarriers
    b += line2_align16(kL2DescRing * sizeof(Line2Desc));                                // descriptors
    b += line2_align16((size_t)kL2DescRing * NBT * line2_frame(B, HP) * 2);             // window masks (u16)
    b += line2_align16((size_t)kL2DescRing * NBT * 4);                                  // walk bounds per block
    b += line2_align16(kL2Stages * (ci ? (4 * P + 8 * LP) * 4 : 0)) + 80 * 4;           // operand words, census table
    b += kL2DescRing * 4;                                                               // claimed runs (producer warps)
    return b;
}

// ---------------------------------------------------------------- cost initialisation of tile groups
// Groups g_first, g_first + g_step, ... < g_end of four positions each; lane q evaluates disparities
// dbase..dbase+3 of the four positions (16 evaluations from 7 operand words, as ci_fill_tile), the argument
// chain of the AD exponential on f32x2 pairs.  Columns 160m-1 / 160m replay the reference's flat indexing
//  in place.
template <bool PLUS, bool FULL_D, int LP>
__device__ __forceinline__ void ci_fill_groups2(float4 *__restrict__ C4, const uint32_t *__restrict__ sOwnP,
                                                const uint32_t *__restrict__ sOwnC, const uint32_t *__restrict__ sOthP,
                                                const uint32_t *__restrict__ sOthC, const float *__restrict__ sLutCen,
                                                float inv_ad, int g_first, int g_step, int g_end, int q, int dbase,
                                                const LineArgs &a, int view, int xb, size_t row)
{
    constexpr int Dc = 4 * LP;
    const int D = a.D, W = a.W;
    bool dv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) dv[j] = dbase + j < D;
    const f32x2_t kMagic = pack2(-8388608.0f, -8388608.0f), kThird = pack2(0.33333333333f, 0.33333333333f),
                  kNegInv = pack2(-inv_ad, -inv_ad), kLog2e = pack2(1.4426950408889634f, 1.4426950408889634f),
                  kNegOne = pack2(-1.0f, -1.0f), kOne = pack2(1.0f, 1.0f);
    // the census table as two halves: Hamming = popc(x) + 32 * bit31(x) (ref_hamdist32, d_alu.cu:7-15) -> the sign of the
    // XOR picks the half, the popcount indexes it
    const char *lutLo = reinterpret_cast<const char *>(sLutCen), *lutHi = lutLo + 128;
    for (int g = g_first; g < g_end; g += g_step) {
        const int p0 = 4 * g;
        const uint4 oP = *reinterpret_cast<const uint4 *>(sOwnP + p0);
        const uint4 oC = *reinterpret_cast<const uint4 *>(sOwnC + p0);
        const int bi = PLUS ? (p0 + 4 * q) : (p0 - 4 * q + Dc - 4);
        const uint4 a0 = *reinterpret_cast<const uint4 *>(sOthP + bi);
        const uint4 a1 = *reinterpret_cast<const uint4 *>(sOthP + bi + 4);
        const uint4 b0 = *reinterpret_cast<const uint4 *>(sOthC + bi);
        const uint4 b1 = *reinterpret_cast<const uint4 *>(sOthC + bi + 4);
        const uint32_t op[4] = {oP.x, oP.y, oP.z, oP.w}, oc[4] = {oC.x, oC.y, oC.z, oC.w};
        const uint32_t wp[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const uint32_t wc[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        // SURVEY Q4 columns (160m-1, 160m) inside this group?  r = first column of the group mod 160 (gx0 >= -HP)
        const int gx0 = xb + p0;
        const int r160 = (gx0 + 2 * kRefBlockW) % kRefBlockW;
        const bool quirk = (r160 == 0 || r160 > kRefBlockW - 5) && gx0 < W;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float r[4];
#pragma unroll
            for (int jj = 0; jj < 4; jj += 2) {
                float cen[2];
                uint32_t fbits[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int j = jj + u;
                    // left view (PLUS): other = R[x + (d - zd)] -> word i + j  (d_ci_ad.cu:133-144);
                    // right view: other = L[x - (d - zd)] -> word 4 + i - j
                    const int w = PLUS ? (i + j) : (4 + i - j);
                    // (float)sad + 2^23 as bits: the sum of absolute byte differences accumulated onto 0x4B000000
                    // (exact for sad < 2^23; the x byte is 0 in both words)
                    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(fbits[u]) : "r"(op[i]), "r"(wp[w]), "r"(0x4B000000u));
                    const uint32_t x = oc[i] ^ wc[w];
                    const char *tb = (int)x < 0 ? lutHi : lutLo;
                    cen[u] = *reinterpret_cast<const float *>(tb + 4 * __popc(x));
                }
                // ad_term (kernels_line.cuh) on the pair: the same five roundings per element, in the same order
                f32x2_t t = add2(pack2(__uint_as_float(fbits[0]), __uint_as_float(fbits[1])), kMagic);
                t = mul2(t, kThird);
                t = mul2(t, kNegInv);
                t = mul2(t, kLog2e);
                float t0, t1, e0, e1;
                unpack2(t, t0, t1);
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(t0));
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(t1));
                f32x2_t c = fma2(pack2(e0, e1), kNegOne, kOne);  // 1 - e, one rounding
                c = add2(c, pack2(cen[0], cen[1]));
                unpack2(c, r[jj], r[jj + 1]);
            }
            const int p = p0 + i;
            if (quirk) {
                const int gx = gx0 + i, tx = gx % kRefBlockW;
                if ((tx == 0 || tx == kRefBlockW - 1) && gx >= 0 && gx < W) {
                    // the reference's 160-wide blocks read one slot outside their half at tx = 0 / 159
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const CiOperands o = ref_ci_operands(view, gx, dbase + j < D ? dbase + j : 0, D, a.zd, W, a.pixL + row,
                                                             a.pixR + row, a.cenL + row, a.cenR + row);
                        const int sad = (int)__vsadu4(o.ad_own, o.ad_other);
                        const int ham = ref_hamdist32(o.cen_own, o.cen_other);
                        r[j] = __fadd_rn(ad_term(sad, inv_ad), sLutCen[ham]);
                    }
                }
            }
            if (!FULL_D) {
#pragma unroll
                for (int j = 0; j < 4; ++j) r[j] = dv[j] ? r[j] : 0.0f;
            }
            C4[(size_t)p * LP + q] = make_float4(r[0], r[1], r[2], r[3]);
        }
    }
}

// ---------------------------------------------------------------- window masks (producer warps)
// Masks of one tile: block b = outputs b*B .. b*B+B-1, frame position k = tile position b*B + k (every window of
// the block lies inside [0, B + 2*HP): arms <= usd <= HP).  Bit i+1 of mask[b][k] = frame position k inside the
// window of output i, [i + HP - A_i, i + HP + B_i).  The kL2Producers warps split the blocks between them; inside
// a warp, 8 lanes share a block and split its frame.  The first lane of a block also writes the block's walk
// bounds (first position | end position << 16; 0 = nothing to add).
constexpr int kL2Producers = 4;

// This lane's output block of a tile of nb blocks (one block per group of LPB lanes: 128 / LPB blocks per tile
// over the four producer warps) or -1
template <int LPB>
__device__ __forceinline__ int mask_lane_block(int pw, int lane, int nb)
{
    const int i = pw * (32 / LPB) + lane / LPB;
    return i < nb ? i : -1;
}

// Masks and walk bounds of one block from its B arm words (0 = output outside the line): this lane writes the
// frame positions of its slice (LPB lanes share a block), the first lane of the block the bounds.
template <int B, bool VERT, int LPB>
__device__ __forceinline__ void masks_from_arms(uint16_t *__restrict__ mrow, uint32_t *__restrict__ bound, const uint32_t ar[B],
                                                int HP, int lane)
{
    const int FR = line2_frame(B, HP), PL = (FR + LPB - 1) / LPB, slice = lane % LPB;
    uint32_t s_rel[B], len[B];
    int first = 0x7fffffff, end = 0;
#pragma unroll
    for (int i = 0; i < B; ++i) {
        const int A = VERT ? arm_up(ar[i]) : arm_left(ar[i]), Bn = VERT ? arm_down(ar[i]) : arm_right(ar[i]);
        s_rel[i] = (uint32_t)(i + HP - A);
        len[i] = (uint32_t)(A + Bn);
        if (A + Bn > 0) {
            first = min(first, i + HP - A);
            end = max(end, i + HP + Bn);
        }
    }
    const int k1 = min(FR, (slice + 1) * PL);
    for (int k = slice * PL; k < k1; ++k) {
        uint32_t m = 0;
#pragma unroll
        for (int i = 0; i < B; ++i) m |= (((uint32_t)k - s_rel[i]) < len[i]) ? (2u << i) : 0u;
        mrow[k] = (uint16_t)m;
    }
    if (slice == 0) *bound = end > 0 ? ((uint32_t)first | ((uint32_t)end << 16)) : 0u;
}

// ---------------------------------------------------------------- the window walk of one output block
// tq: shared address of this lane's float4 of tile position 0; block outputs o0..o0+B-1; mrow: shared address
// of the block's masks; bd: its walk bounds.
template <int B, uint32_t PBY = kL2PosBytes>
__device__ __forceinline__ void sum_block_span(uint32_t tq, uint32_t mrow, uint32_t first, uint32_t end, int o0, float4 acc[B])
{
#pragma unroll
    for (int i = 0; i < B; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (first >= end) return;
    uint32_t p = tq + ((uint32_t)o0 + first) * PBY, mp = mrow + 2u * first;
    const uint32_t n = end - first;
#pragma unroll kL2Unroll
    for (uint32_t k = 0; k < n; ++k, p += PBY, mp += 2) {
        const float4 v = lds128<0>(p);
        const uint32_t m = lds16(mp);
#pragma unroll
        for (int i = 0; i < B; ++i)
            if (m & (2u << i)) acc4(acc[i], v);
    }
}
template <int B>
__device__ __forceinline__ void sum_block_masked(uint32_t tq, uint32_t mrow, uint32_t bd, int o0, float4 acc[B])
{
    sum_block_span<B, kL2PosBytes>(tq, mrow, bd & 0xffffu, bd >> 16, o0, acc);
}

// ---------------------------------------------------------------- winner-takes-all of one output block
// dc_wta_kernel (d_dc_wta.cu:9-35): strict '>' from FLT_MAX, first minimum wins.  Aggregated ADCensus costs are
// >= +0, so their bit patterns order like the floats: lane minimum over its four disparities, CREDUX.MIN across
// the warp, then CREDUX.MIN over the disparity indices of the lanes that hold the minimum.  Lane i keeps output
// i's result; one store (or one atomicMin on (cost, d) keys when the disparity range spans several chunks).
template <int B, bool FULL_D, int LP = 32>
__device__ __forceinline__ void wta_block(const float4 acc[B], const LineArgs &a, int vslot, int ln, int x0, int nvalid, int d0,
                                          int lane_in)
{
    // a pixel is the LP lanes of one sub-warp: the reductions run over that sub-warp only
    const int lane = lane_in % LP;
    const unsigned team = LP == 32 ? 0xffffffffu : (((1u << LP) - 1u) << ((lane_in / LP) * LP));
    const int dq = d0 + 4 * lane;
    uint32_t my_d = 0, my_m = 0;
#pragma unroll
    for (int i = 0; i < B; ++i) {
        uint32_t bx = __float_as_uint(acc[i].x), by = __float_as_uint(acc[i].y), bz = __float_as_uint(acc[i].z),
                 bw = __float_as_uint(acc[i].w);
        if (!FULL_D) {  // padded disparities never win
            if (dq + 0 >= a.D) bx = 0xffffffffu;
            if (dq + 1 >= a.D) by = 0xffffffffu;
            if (dq + 2 >= a.D) bz = 0xffffffffu;
            if (dq + 3 >= a.D) bw = 0xffffffffu;
        }
        const uint32_t lm = min(min(bx, by), min(bz, bw));
        const uint32_t mn = __reduce_min_sync(team, lm);
        const uint32_t j = bx == mn ? 0u : (by == mn ? 1u : (bz == mn ? 2u : 3u));
        const uint32_t cand = lm == mn ? (uint32_t)dq + j : 0x7fffffffu;
        const uint32_t d = __reduce_min_sync(team, cand);
        if (lane == i) { my_d = d; my_m = mn; }
    }
    if (lane < nvalid) {
        const size_t pix = (size_t)ln * a.W + (x0 + lane);
        if (!a.use_keys) {
            a.disp[vslot][pix] = (float)(int)my_d - (float)a.zd;
        } else {
            const unsigned long long key = ((unsigned long long)float_orderable(__uint_as_float(my_m)) << 32) | my_d;
            atomicMin(a.wta_key[vslot] + pix, key);
        }
    }
}

// ---------------------------------------------------------------- the kernel
// LP = float4 lanes per pixel (32: one warp per pixel, 128 disparities per chunk; 16: two pixels per warp, num_disp <= 64).
// With LP < 32 the sub-warps of a warp work on neighbouring blocks of the same tile, each lane with its own block's
// masks, under one loop that spans the union of their walks.
template <int MODE, int NW, int B, int NS, int LP = 32>
__global__ void __launch_bounds__((NW + kL2Producers) * 32, 1)
k_line2(const __grid_constant__ Line2Args L, const __grid_constant__ CUtensorMap tmap)
{
    extern __shared__ __align__(16) unsigned char smem_l2[];
    constexpr bool VERT = (MODE == LM_V);
    constexpr bool CI = (MODE == LM_CI_H);
    const LineArgs &a = L.a;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int SUB = 32 / LP, Dc = 4 * LP;      // pixels per warp, disparities per chunk
    constexpr uint32_t PBY = 16u * LP;              // bytes per tile position
    const int sub = lane / LP, q = lane % LP;
    const int S = L.S, HP = L.HP, P = L.P, W = a.W;

    // ---- shared-memory carve-up (line2_smem_bytes); the tiles start on a 128-byte boundary (tensor copies)
    unsigned char *smem_raw = smem_l2 + ((128u - (smem_u32(smem_l2) & 127u)) & 127u);
    const uint32_t tile_bytes = (uint32_t)P * PBY;
    unsigned char *sp = smem_raw + (size_t)NS * tile_bytes;
    const uint32_t bars = smem_u32(sp);  // [0..NS) full, then empty, fullO, emptyO (NS <= 4)
    sp += 128;
    Line2Desc *desc = reinterpret_cast<Line2Desc *>(sp);
    sp += line2_align16(kL2DescRing * sizeof(Line2Desc));
    const int FR = line2_frame(B, HP), NBT = (S + B - 1) / B;
    constexpr int LPB = (B <= 4 || LP < 32) ? 4 : 8;  // producer lanes per block: up to 32 or 16 blocks per tile
    uint16_t *sMask = reinterpret_cast<uint16_t *>(sp);  // [ring][block][frame position]
    sp += line2_align16((size_t)kL2DescRing * NBT * FR * 2);
    uint32_t *sBounds = reinterpret_cast<uint32_t *>(sp);  // [ring][block]
    sp += line2_align16((size_t)kL2DescRing * NBT * 4);
    uint32_t *sOps = reinterpret_cast<uint32_t *>(sp);  // CI: per stage ownP[P] ownC[P] othP[P+128] othC[P+128]
    const int OPS = CI ? 4 * P + 2 * Dc : 0;
    float *sLutCen = reinterpret_cast<float *>(sp + line2_align16((size_t)NS * OPS * 4));

    auto full = [&](int s) { return bars + 8u * s; };
    auto empty = [&](int s) { return bars + 8u * (NS + s); };
    auto fullO = [&](int s) { return bars + 8u * (2 * NS + s); };
    auto emptyO = [&](int s) { return bars + 8u * (3 * NS + s); };

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) {
            // CI: the consumer warps complete a tile; else every producer lane (masks written) + the copy's bytes
            mbar_init(full(s), CI ? NW : 32 * kL2Producers + 1);
            mbar_init(empty(s), NW);
            mbar_init(fullO(s), 64 * kL2Producers);  // per producer lane: its cp.async landed + its masks written
            mbar_init(emptyO(s), NW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (CI && tid < kCenLutSize) sLutCen[tid] = a.lutCen[tid];
    __syncthreads();

    const size_t pos_stride4 = VERT ? (size_t)W * a.LPtot : (size_t)a.LPtot;  // float4 between line positions
    const int LEN = VERT ? a.v_end : W;
    const int LINE0 = VERT ? a.v_begin : 0;
    const int IN_LO = VERT ? a.v_lo : 0, IN_HI = VERT ? a.v_hi : W;

    // =========================================================== producer warps
    // All of them walk the same tile sequence (every warp claims through its own shared-memory copy of the run
    // the leader fetched); the leader (first producer warp) also writes the descriptor and starts the copy.
    if (warp >= NW) {
        const int pw = warp - NW;
        volatile int *sClaim = reinterpret_cast<volatile int *>(sp + line2_align16((size_t)NS * OPS * 4) + 80 * 4);  // [kL2DescRing]
        int seg = 0, seg_end = 0, ln = 0, vslot = 0, chunk = 0, first = 0, nclaim = 0;
        uint32_t cur[B], pend_bar = 0;
        int pend_ring = 0, pend_b = -1;
#pragma unroll
        for (int i = 0; i < B; ++i) cur[i] = 0u;
        for (int m = 0;; ++m) {
            const int st = m % NS, ring = m % kL2DescRing;
            if (m >= NS) mbar_wait(CI ? emptyO(st) : empty(st), ((m / NS) - 1) & 1);
            int valid = 1;
            if (seg == seg_end) {  // next run of consecutive segments
                // the leader draws it from the global counter and hands it to the other producer warps
                int c = 0;
                if (pw == 0) {
                    if (lane == 0) {
                        c = atomicAdd(L.counter, 1);
                        sClaim[nclaim % kL2DescRing] = c;
                    }
                    __syncwarp();
                }
                asm volatile("bar.sync 1, %0;" ::"n"(32 * kL2Producers) : "memory");
                c = sClaim[nclaim % kL2DescRing];
                ++nclaim;
                if (c >= L.nclaims) {
                    valid = 0;
                } else {
                    const int per_z = L.nlines * L.claims_per_line;
                    const int z = c / per_z, r = c - z * per_z;
                    const int li = r / L.claims_per_line, k = r - li * L.claims_per_line;
                    ln = a.ln_first + li;
                    vslot = z / a.nchunks;
                    chunk = z - vslot * a.nchunks;
                    seg = k * L.tiles_per_claim;
                    seg_end = min(seg + L.tiles_per_claim, L.tiles_per_line);
                    first = 1;
                }
            }
            const int t0 = LINE0 + seg * S;
            const int Sact = valid ? min(S, LEN - t0) : 0;
            if (pw == 0 && lane == 0) {
                Line2Desc d;
                d.valid = valid; d.ln = ln; d.seg = seg; d.vslot = vslot; d.chunk = chunk; d.t0 = t0; d.Sact = Sact; d.first = first;
                desc[ring] = d;
            }
            const uint32_t bar = CI ? fullO(st) : full(st);
            if (!CI && pw == 0 && lane == 0) {
                // tile <- volume.  Positions outside the readable part of the line are never inside a window.
                if (!valid) {
                    mbar_arrive_expect_tx(bar, 0u);
                } else if (L.use_tmap) {
                    mbar_arrive_expect_tx(bar, tile_bytes);
                    const uint32_t dst = smem_u32(smem_raw) + (uint32_t)st * tile_bytes;
                    // several copies per tile: more requests of the copy engine in flight
                    for (int p0 = 0; p0 < P; p0 += L.tmap_rows) {
                        if (VERT) tma_load_4d(dst + (uint32_t)p0 * PBY, &tmap, chunk * Dc, ln, t0 - HP - L.row_bias + p0, vslot, bar);
                        else tma_load_4d(dst + (uint32_t)p0 * PBY, &tmap, chunk * Dc, t0 - HP + p0, ln - L.row_bias, vslot, bar);
                    }
                }
            }
            if (!CI && valid && !L.use_tmap && pw == 0) {
                const int p_lo = max(0, IN_LO + HP - t0), p_hi = min(P, IN_HI - t0 + HP);
                const int np = max(p_hi - p_lo, 0);
                if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)np * PBY);
                const size_t line_base4 = (VERT ? (size_t)ln * a.LPtot : (size_t)ln * W * a.LPtot) + (size_t)chunk * LP;
                const char *src0 = reinterpret_cast<const char *>(a.in[vslot] + line_base4) +
                                   (long long)(t0 - HP + p_lo) * (long long)pos_stride4 * 16;
                const uint32_t dst0 = smem_u32(smem_raw) + (uint32_t)st * tile_bytes + (uint32_t)p_lo * PBY;
                const long long gs = (long long)pos_stride4 * 16;
                for (int p = lane; p < np; p += 32)
                    bulk_g2s(dst0 + (uint32_t)p * PBY, src0 + (long long)p * gs, PBY, bar);
            }
            if (valid && CI) {
                const int view = a.view_first + vslot;
                const int d0 = a.d_first + chunk * Dc;
                const size_t row = (size_t)ln * W;
                const int xb = t0 - HP;
                // other-view words start at the column that makes every thread's 8-word window 16-byte aligned
                const int xo = (view == 0) ? (xb - a.zd + d0) : (xb + a.zd - d0 - Dc);
                const uint32_t *gOwnP = (view == 0 ? a.pixL : a.pixR) + row, *gOwnC = (view == 0 ? a.cenL : a.cenR) + row;
                const uint32_t *gOthP = (view == 0 ? a.pixR : a.pixL) + row, *gOthC = (view == 0 ? a.cenR : a.cenL) + row;
                const uint32_t o0 = smem_u32(sOps + (size_t)st * OPS);
                // a continued run only needs the operands of its S new positions; the first tile all of them
                const int i0 = first ? 0 : 2 * HP;
                const int pl = pw * 32 + lane;
                for (int i = i0 + pl; i < P; i += 32 * kL2Producers) {
                    const int x = clampi(xb + i, 0, W - 1);
                    cp_async4(o0 + 4u * i, gOwnP + x, 4u);
                    cp_async4(o0 + 4u * (P + i), gOwnC + x, 4u);
                }
                for (int i = i0 + pl; i < P + Dc; i += 32 * kL2Producers) {
                    const int x = clampi(xo + i, 0, W - 1);
                    cp_async4(o0 + 4u * (2 * P + i), gOthP + x, 4u);
                    cp_async4(o0 + 4u * (3 * P + Dc + i), gOthC + x, 4u);
                }
            }
            if (CI) cp_async_arrive_noinc(bar);
            // Window masks, one tile behind: the arm words of THIS tile are requested now and turned into masks in
            // the next trip, so their load latency never sits between two tiles.
            uint32_t nxt[B];
            const int my_b = valid ? mask_lane_block<LPB>(pw, lane, NBT) : -1;
#pragma unroll
            for (int i = 0; i < B; ++i) {
                const int o = my_b * B + i;
                nxt[i] = 0u;
                if (my_b >= 0 && o < Sact) nxt[i] = __ldg(a.arms[vslot] + (VERT ? (size_t)(t0 + o) * W + ln : (size_t)ln * W + (t0 + o)));
            }
            if (pend_bar) {
                if (pend_b >= 0)
                    masks_from_arms<B, VERT, LPB>(sMask + ((size_t)pend_ring * NBT + pend_b) * FR, sBounds + pend_ring * NBT + pend_b, cur, HP,
                                                  lane);
                mbar_arrive(pend_bar);  // releases this lane's masks (leader lane 0: and the descriptor)
            }
#pragma unroll
            for (int i = 0; i < B; ++i) cur[i] = nxt[i];
            pend_bar = bar; pend_ring = ring; pend_b = my_b;
            if (!valid) {
                mbar_arrive(bar);
                break;
            }
            ++seg;
            first = 0;
        }
        return;
    }

    // =========================================================== consumer warps
    const uint32_t tq0 = smem_u32(smem_raw) + (uint32_t)q * 16u;
    const long long ostride = (long long)pos_stride4 * 16;
    float4 *const C4base = reinterpret_cast<float4 *>(smem_raw);

    auto ci_fill = [&](int m, const Line2Desc &d, bool carry, int prev_stage) {
        // consumer-side production of tile m (pass 1): operands -> 128 costs per position
        const int st = m % NS;
        const int view = a.view_first + d.vslot;
        const int d0 = a.d_first + d.chunk * Dc;
        const uint32_t *ops = sOps + (size_t)st * OPS;
        float4 *C4 = C4base + (size_t)st * P * LP;
        int g_begin = 0;
        if (carry) {
            // positions [0, 2*HP) of this tile are positions [S, S + 2*HP) of the previous one
            const float4 *src = C4base + (size_t)prev_stage * P * LP + (size_t)S * LP;
            for (int p = warp * SUB + sub; p < 2 * HP; p += NW * SUB) C4[(size_t)p * LP + q] = src[(size_t)p * LP + q];
            g_begin = (2 * HP) / 4;
        }
        const bool full_d = d0 + Dc <= a.D;
        const size_t row = (size_t)d.ln * W;
        const int xb = d.t0 - HP;
        // rotate the first group among the warps so that a longer first tile does not always load warp 0
        const int g_first = g_begin + ((warp + m) % NW) * SUB + sub;
#define S2MV_L2_FILL(PLUS, FULL)                                                                                             \
    ci_fill_groups2<PLUS, FULL, LP>(C4, ops, ops + P, ops + 2 * P, ops + 3 * P + Dc, sLutCen, a.inv_ad, g_first, NW * SUB, P / 4, \
                                    q, d0 + 4 * q, a, view, xb, row)
        if (view == 0) { if (full_d) S2MV_L2_FILL(true, true); else S2MV_L2_FILL(true, false); }
        else           { if (full_d) S2MV_L2_FILL(false, true); else S2MV_L2_FILL(false, false); }
#undef S2MV_L2_FILL
    };

    if (CI) {
        // prologue: tile 0 is produced before the loop
        mbar_wait(fullO(0), 0);
        const Line2Desc d = desc[0];
        if (d.valid) ci_fill(0, d, false, 0);
        __syncwarp();
        if (lane == 0) { mbar_arrive(full(0)); mbar_arrive(emptyO(0)); }
    }

    for (int n = 0;; ++n) {
        const int st = n % NS, ring = n % kL2DescRing;
        mbar_wait(full(st), (n / NS) & 1);
        const Line2Desc d = desc[ring];
        if (!d.valid) break;
        if (CI) {
            const int m = n + 1, st1 = m % NS;
            mbar_wait(fullO(st1), (m / NS) & 1);
            const Line2Desc d1 = desc[m % kL2DescRing];
            if (d1.valid) ci_fill(m, d1, !d1.first, st);
            __syncwarp();
            if (lane == 0) { mbar_arrive(full(st1)); mbar_arrive(emptyO(st1)); }
        }

        const int vslot = d.vslot, chunk = d.chunk, ln = d.ln, t0 = d.t0, Sact = d.Sact;
        const int d0 = a.d_first + chunk * Dc;
        const bool full_d = d0 + Dc <= a.D;
        const uint32_t tq = tq0 + (uint32_t)st * tile_bytes;
        const size_t line_base4 = (VERT ? (size_t)ln * a.LPtot : (size_t)ln * W * a.LPtot) + (size_t)chunk * LP + q;
        const int nblocks = (Sact + B - 1) / B;
        const bool has_peer = MODE != LM_H_WTA && (a.peer_out[0][vslot] != nullptr || a.peer_out[1][vslot] != nullptr);
        for (int b0 = warp * SUB; b0 < nblocks; b0 += NW * SUB) {
            // this lane's block (every sub-warp its own; NBT is a multiple of SUB, blocks past the line end have empty masks)
            const int b = b0 + sub;
            const int o0 = b * B, nvalid = max(0, min(B, Sact - o0));
            float4 acc[B];
            const uint32_t bd = sBounds[ring * NBT + b];
            uint32_t first = bd ? (bd & 0xffffu) : 0xffffu, end = bd >> 16;
            if (SUB > 1) {  // one loop over the union of the sub-warps' walks; a lane's masks are 0 outside its own
                first = __reduce_min_sync(0xffffffffu, first);
                end = __reduce_max_sync(0xffffffffu, end);
            }
            sum_block_span<B, PBY>(tq, smem_u32(sMask + ((size_t)ring * NBT + b) * FR), first, end, o0, acc);
            if (MODE != LM_H_WTA) {
                char *dstp = reinterpret_cast<char *>(a.out[vslot] + line_base4) + (long long)(t0 + o0) * ostride;
#pragma unroll
                for (int i = 0; i < B; ++i)
                    if (i < nvalid) *reinterpret_cast<float4 *>(dstp + i * ostride) = acc[i];
                if (has_peer) {  // this launch feeds a neighbouring band's halo rows (see k_line)
                    const bool peer_up = a.peer_out[0][vslot] != nullptr, peer_dn = a.peer_out[1][vslot] != nullptr;
                    const long long d_up = reinterpret_cast<char *>(a.peer_out[0][vslot]) - reinterpret_cast<char *>(a.out[vslot]);
                    const long long d_dn = reinterpret_cast<char *>(a.peer_out[1][vslot]) - reinterpret_cast<char *>(a.out[vslot]);
#pragma unroll
                    for (int i = 0; i < B; ++i) {
                        const int r = VERT ? (t0 + o0 + i) : ln;  // local image row of output i
                        if (i < nvalid && r < a.peer_lo_end && peer_up) *reinterpret_cast<float4 *>(dstp + i * ostride + d_up) = acc[i];
                        if (i < nvalid && r >= a.peer_hi_begin && peer_dn) *reinterpret_cast<float4 *>(dstp + i * ostride + d_dn) = acc[i];
                    }
                }
            } else {
                if (full_d) wta_block<B, true, LP>(acc, a, vslot, ln, t0 + o0, nvalid, d0, lane);
                else wta_block<B, false, LP>(acc, a, vslot, ln, t0 + o0, nvalid, d0, lane);
            }
        }
        if (!CI) {
            __syncwarp();
            if (lane == 0) mbar_arrive(empty(st));
        }
    }
}


// =====================================================================================================
// k_line_vv — the two vertical passes (V2 and V3 of H, V, V, H; d_ca_cross.cu:261-264) in ONE kernel: the
// volume between them never exists in HBM (the separate passes move 4V per view for it, this one 2V).
//
// Both passes sum along the same column with the SAME windows (same pixel, same up/down arms):
//     V2[r] = sum_{k in [r-U_r, r+D_r)} H1[k]        V3[r] = sum_{k in [r-U_r, r+D_r)} V2[k]
// A CTA walks down a column in tiles of S rows.  Step j:
//   * the producer warps fetch H1 rows [jS, jS + S + 2u) with one tensor copy (ring of three stages) and build
//     the window masks of the row blocks that become computable (each row block's masks serve both passes);
//   * the A warps (one block of B rows each) compute V2 for the S NEW rows [jS+u, jS+S+u) into the V2 buffer of
//     the step, after copying the 2u rows that buffer shares with the previous step's (shared memory to shared
//     memory: no V2 row is ever computed twice); at the top of a column they also compute rows [0, u);
//   * the B warps compute V3 for rows [jS, jS+S) from that V2 buffer (rows [jS-u, jS+S+u)) and store them.
// A runs one step ahead of B (two V2 buffers); "V2 of step n complete" and "V3 of step n complete" are
// mbarriers.  Every accumulator receives exactly the reference's additions in the reference's order
// (d_ca_cross_sum.cu:148-198), so the result equals the two separate passes bit for bit.
struct LineVVArgs {
    LineArgs a;           // in = H1 volume (unused when the tensor map is), out = V3 volume, arms
    int S, HP, P;         // rows per tile, halo rows (>= usd, a multiple of B), S + 2*HP
    int tiles_per_col, ncols, nz, nclaims;
    int *counter;
    int tmap_rows;        // rows per tensor copy (the map's box): a tile is P / tmap_rows copies
    // Rows stored, [out_r0, out_r1): all a.H rows of a whole image.  A row band runs the kernel over the rows its buffer A
    // holds (its own rows and 2*usd rows of each neighbour) as if they were an image: arms are cut at the first and last
    // of them, which changes V2 only within usd rows of those edges and V3 only within 2*usd -- outside the own rows
    // [out_r0, out_r1), the only ones stored.
    int out_r0, out_r1;
};

// arm word of row r of an H-row column with the vertical arms cut at rows 0 and H-1 (whole images: already are)
__device__ __forceinline__ uint32_t arm_cut_v(uint32_t w, int r, int H)
{
    const uint32_t up = min(w & 0xffu, (uint32_t)r), dn = min((w >> 8) & 0xffu, (uint32_t)(H - 1 - r));
    return (w & 0xffff0000u) | (dn << 8) | up;
}

struct LineVVDesc {
    int valid, col, vslot, chunk, j, colbase, pad0, pad1;
};

constexpr int kVVMaskRing = 64;  // row blocks whose masks are kept

template <int NA, int B>
__host__ __device__ inline size_t linevv_smem_bytes(int S, int HP)
{
    const size_t P = (size_t)S + 2 * HP;
    size_t b = 128 + 5 * P * kL2PosBytes;                                         // 3 H1 stages + 2 V2 buffers
    b += 128;                                                                     // mbarriers
    b += line2_align16(kL2DescRing * sizeof(LineVVDesc));
    b += line2_align16((size_t)kVVMaskRing * line2_frame(B, HP) * 2);             // masks
    b += line2_align16((size_t)kVVMaskRing * 4);                                  // bounds
    b += kL2DescRing * 4;                                                         // claimed columns
    return b;
}

template <int NA, int B>
__global__ void __launch_bounds__((2 * NA + kL2Producers) * 32, 1)
k_line_vv(const __grid_constant__ LineVVArgs L, const __grid_constant__ CUtensorMap tmap)
{
    extern __shared__ __align__(16) unsigned char smem_l2[];
    const LineArgs &a = L.a;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int S = L.S, HP = L.HP, P = L.P, W = a.W, H = a.H;
    const int NBS = S / B, NBH = HP / B;  // row blocks per tile / per halo

    unsigned char *smem_raw = smem_l2 + ((128u - (smem_u32(smem_l2) & 127u)) & 127u);
    const uint32_t tile_bytes = (uint32_t)P * kL2PosBytes;
    const uint32_t h1_base = smem_u32(smem_raw), v2_base = h1_base + 3u * tile_bytes;
    unsigned char *sp = smem_raw + 5 * (size_t)tile_bytes;
    const uint32_t bars = smem_u32(sp);  // fullH[3] emptyH[3] v2done[2] v3done[2]
    sp += 128;
    LineVVDesc *desc = reinterpret_cast<LineVVDesc *>(sp);
    sp += line2_align16(kL2DescRing * sizeof(LineVVDesc));
    const int FR = line2_frame(B, HP);
    uint16_t *sMask = reinterpret_cast<uint16_t *>(sp);
    sp += line2_align16((size_t)kVVMaskRing * FR * 2);
    uint32_t *sBounds = reinterpret_cast<uint32_t *>(sp);
    sp += line2_align16((size_t)kVVMaskRing * 4);
    volatile int *sClaim = reinterpret_cast<volatile int *>(sp);

    auto fullH = [&](int s) { return bars + 8u * s; };
    auto emptyH = [&](int s) { return bars + 8u * (3 + s); };
    auto v2done = [&](int s) { return bars + 8u * (6 + s); };
    auto v3done = [&](int s) { return bars + 8u * (8 + s); };

    if (tid == 0) {
        for (int s = 0; s < 3; ++s) {
            mbar_init(fullH(s), 32 * kL2Producers + 1);
            mbar_init(emptyH(s), NA);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(v2done(s), NA);
            mbar_init(v3done(s), NA);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // =========================================================== producer warps
    if (warp >= 2 * NA) {
        const int pw = warp - 2 * NA;
        int j = 0, jend = 0, col = 0, vslot = 0, chunk = 0, nclaim = 0, colbase = 0, nextbase = 0;
        uint32_t cur[B], pend_bar = 0;
        int pend_slot = -1, pend_r0 = 0;
#pragma unroll
        for (int k = 0; k < B; ++k) cur[k] = 0u;
        for (int n = 0;; ++n) {
            const int st = n % 3, ring = n % kL2DescRing;
            if (n >= 3) mbar_wait(emptyH(st), ((n / 3) - 1) & 1);
            int valid = 1;
            if (j == jend) {  // next column
                int c = 0;
                if (pw == 0) {
                    if (lane == 0) {
                        c = atomicAdd(L.counter, 1);
                        sClaim[nclaim % kL2DescRing] = c;
                    }
                    __syncwarp();
                }
                asm volatile("bar.sync 1, %0;" ::"n"(32 * kL2Producers) : "memory");
                c = sClaim[nclaim % kL2DescRing];
                ++nclaim;
                if (c >= L.nclaims) {
                    valid = 0;
                } else {
                    const int z = c / L.ncols;
                    col = c - z * L.ncols;
                    vslot = z / a.nchunks;
                    chunk = z - vslot * a.nchunks;
                    j = 0;
                    jend = L.tiles_per_col;
                    colbase = nextbase;
                    nextbase = (nextbase + NBH + NBS * L.tiles_per_col) % kVVMaskRing;
                }
            }
            if (pw == 0 && lane == 0) {
                LineVVDesc d;
                d.valid = valid; d.col = col; d.vslot = vslot; d.chunk = chunk; d.j = j; d.colbase = colbase; d.pad0 = d.pad1 = 0;
                desc[ring] = d;
                if (valid) {
                    mbar_arrive_expect_tx(fullH(st), tile_bytes);
                    for (int p0 = 0; p0 < P; p0 += L.tmap_rows)
                        tma_load_4d(h1_base + (uint32_t)st * tile_bytes + (uint32_t)p0 * kL2PosBytes, &tmap, chunk * 128, col, j * S + p0, vslot,
                                    fullH(st));
                } else {
                    mbar_arrive_expect_tx(fullH(st), 0u);
                }
            }
            // Masks of the row blocks that become computable at this step: g = NBH + j*NBS + i (rows [jS + HP + B*i, +B)),
            // i < NBS; at the top of a column also g = 0..NBH-1 (rows [0, HP)).  One block per group of 8 lanes; the arm
            // words are requested now and turned into masks in the next trip (their latency stays off the tile cadence).
            uint32_t nxt[B];
            int my_slot = -1, my_r0 = 0;
            {
                const int nb = valid ? NBS + (j == 0 ? NBH : 0) : 0;
                const int i = mask_lane_block<8>(pw, lane, nb);
                const int g = (j == 0 ? 0 : NBH + j * NBS) + i;
                if (i >= 0) my_slot = (colbase + g) % kVVMaskRing;
                my_r0 = g * B;
#pragma unroll
                for (int k = 0; k < B; ++k) {
                    const int r = g * B + k;
                    nxt[k] = 0u;
                    if (i >= 0 && r < H) nxt[k] = __ldg(a.arms[vslot] + (size_t)r * W + col);
                }
            }
            if (pend_bar) {
                if (pend_slot >= 0) {
                    // (the cut is applied here, a trip after the request, not on the fresh load: the request's latency
                    // must stay off this loop)
#pragma unroll
                    for (int k = 0; k < B; ++k) cur[k] = arm_cut_v(cur[k], pend_r0 + k, H);
                    masks_from_arms<B, true, 8>(sMask + (size_t)pend_slot * FR, sBounds + pend_slot, cur, HP, lane);
                }
                mbar_arrive(pend_bar);
            }
#pragma unroll
            for (int k = 0; k < B; ++k) cur[k] = nxt[k];
            pend_bar = fullH(st); pend_slot = my_slot; pend_r0 = my_r0;
            if (!valid) {
                mbar_arrive(fullH(st));
                break;
            }
            ++j;
        }
        return;
    }

    const bool isA = warp < NA;
    const int w = isA ? warp : warp - NA;
    const uint32_t lane16 = (uint32_t)lane * 16u;

    if (isA) {
        // =========================================================== A warps: V2 = V(H1)
        for (int n = 0;; ++n) {
            const int st = n % 3, vb = n & 1;
            mbar_wait(fullH(st), (n / 3) & 1);
            const LineVVDesc d = desc[n % kL2DescRing];
            if (n >= 1) mbar_wait(v2done(vb ^ 1), ((n - 1) >> 1) & 1);  // the previous step's V2 rows (carry source)
            if (n >= 2) mbar_wait(v3done(vb), ((n - 2) >> 1) & 1);      // this V2 buffer's last readers
            if (!d.valid) {
                __syncwarp();
                if (lane == 0) mbar_arrive(v2done(vb));
                break;
            }
            const uint32_t h1 = h1_base + (uint32_t)st * tile_bytes + lane16;     // position 0 = row jS
            const uint32_t v2 = v2_base + (uint32_t)vb * tile_bytes + lane16;     // position 0 = row jS - HP
            if (d.j > 0) {
                // rows [jS - HP, jS + HP) were computed by the previous step: its positions [S, S + 2*HP)
                const uint32_t src = v2_base + (uint32_t)(vb ^ 1) * tile_bytes + lane16 + (uint32_t)S * kL2PosBytes;
                for (int p = w; p < 2 * HP; p += NA) {
                    const float4 v = lds128<0>(src + (uint32_t)p * kL2PosBytes);
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(v2 + (uint32_t)p * kL2PosBytes), "f"(v.x), "f"(v.y),
                                 "f"(v.z), "f"(v.w)
                                 : "memory");
                }
            }
            // new rows: blocks i < NBS are rows [jS + HP + B*i, +B) = row blocks g = NBH + j*NBS + i; at the top
            // of a column blocks NBS.. are rows [0, HP) = g = 0..NBH-1
            const int nb = NBS + (d.j == 0 ? NBH : 0);
            for (int i = w; i < nb; i += NA) {
                const int g = i < NBS ? NBH + d.j * NBS + i : i - NBS;
                const int slot = (d.colbase + g) % kVVMaskRing;
                const int r0 = g * B;                       // first row of the block
                if (r0 + B <= L.out_r0 - HP || r0 >= L.out_r1 + HP) continue;  // no stored row reads this block
                // frame position 0 = row r0 - HP = H1 tile position r0 - HP - jS
                float4 acc[B];
                sum_block_masked<B>(h1, smem_u32(sMask + (size_t)slot * FR), sBounds[slot], r0 - HP - d.j * S, acc);
                const uint32_t dst = v2 + (uint32_t)(r0 - d.j * S + HP) * kL2PosBytes;  // V2 buffer position of row r0
#pragma unroll
                for (int k = 0; k < B; ++k)
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(dst + (uint32_t)k * kL2PosBytes), "f"(acc[k].x),
                                 "f"(acc[k].y), "f"(acc[k].z), "f"(acc[k].w)
                                 : "memory");
            }
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(emptyH(st));
                mbar_arrive(v2done(vb));
            }
        }
    } else {
        // =========================================================== B warps: V3 = V(V2), stored
        const long long ostride = (long long)W * a.LPtot * 16;  // bytes between rows of a column
        for (int n = 0;; ++n) {
            const int vb = n & 1;
            mbar_wait(v2done(vb), (n >> 1) & 1);
            const LineVVDesc d = desc[n % kL2DescRing];
            if (!d.valid) break;
            const uint32_t v2 = v2_base + (uint32_t)vb * tile_bytes + lane16;  // position 0 = row jS - HP
            char *colp = reinterpret_cast<char *>(a.out[d.vslot] + (size_t)d.col * a.LPtot + (size_t)d.chunk * 32 + lane);
            for (int i = w; i < NBS; i += NA) {
                const int g = d.j * NBS + i, r0 = g * B;
                if (r0 >= L.out_r1) break;
                if (r0 + B <= L.out_r0) continue;
                const int slot = (d.colbase + g) % kVVMaskRing;
                // frame position 0 = row r0 - HP = V2 buffer position r0 - jS
                float4 acc[B];
                sum_block_masked<B>(v2, smem_u32(sMask + (size_t)slot * FR), sBounds[slot], r0 - d.j * S, acc);
                char *dstp = colp + (long long)r0 * ostride;
#pragma unroll
                for (int k = 0; k < B; ++k)
                    if (r0 + k >= L.out_r0 && r0 + k < L.out_r1) *reinterpret_cast<float4 *>(dstp + k * ostride) = acc[k];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(v3done(vb));
        }
    }
}

}  // namespace s2mv
