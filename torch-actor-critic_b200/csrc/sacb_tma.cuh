// sacb_tma.cuh -- precision 3 with bf16 shadow operands: a TMA-fed tcgen05 GEMM tile.
// Included by sacb_device.cuh after sacb_tcgen05.cuh (inside namespace sacb).
//
//   C[128 x 64] (+epilogue) = sum_r A16(i,r) * B16(j,r), bf16 operands, fp32 accumulation in TMEM.
//
// Every GEMM operand of the phase programs also exists as a bf16 matrix (RunArgs.sh, written by whoever writes the fp32
// original), described by a CUtensorMap the host encoded (sacb_engine.cu: build_tma).  One thread issues
// cp.async.bulk.tensor.2d (SASS UTMALDG) per 64-wide K chunk: the TMA drops a [rows x 64] bf16 box (128-byte rows,
// SWIZZLE_128B) straight into the canonical UMMA shared-memory layout, for reduction-contiguous operands (K-major) and
// for output-contiguous ones (MN-major) alike -- no staging copy, no conversion pass, no per-thread address math.  A
// second thread issues the tcgen05.mma's of a chunk as soon as its `full` mbarrier flips and commits them to the chunk's
// `empty` mbarrier, which hands the stage back to the TMA thread: a 6-stage ring (24 KB per stage) that keeps ~144 KB of
// loads in flight per SM.
//     K-major  box {64 k, rows}: row r at r*128 B; 8-row groups 1024 B apart (SBO); a K=16 MMA step advances 32 B
//     MN-major box {64 mn, 64 k}: k-row at k*128 B; 8-k groups 1024 B apart (SBO); the second 64-wide half of a 128-row A
//              tile is a second box 8192 B further (LBO); a K=16 MMA step advances 2048 B
// (cute/atom/mma_traits_sm100.hpp, make_umma_desc: LayoutType::B128, both majors.)
#pragma once

constexpr int kTmStages = 6;
constexpr int kTmABytes = kBfM * kBfK * 2, kTmBBytes = kBfN * kBfK * 2;      // 16,384 + 8,192
constexpr int kTmStageBytes = kTmABytes + kTmBBytes;
constexpr int kTmMaskOff = kTmStages * kTmStageBytes;                        // 147,456: relu-mask tile [128][68] fp32
constexpr int kTmEpiOff = kTmMaskOff + kBfM * kBfPitchRC * 4;                // 182,272: bias[64], w2[64]
constexpr int kTmBarOff = kTmEpiOff + 2 * kBfN * 4;                          // 182,784: full[6], empty[6], done
constexpr int kTmEnd = kTmBarOff + (2 * kTmStages + 1) * 8;
static_assert(kTmEnd + 1024 <= kT5PipeBytes, "TMA tile buffers (+ 1 KB alignment slack) must fit in the tcgen05 kernel's shared memory");
static_assert(kBfM * kBfPitchRC * 4 <= 2 * kTmStageBytes, "the epilogue's tile buffer aliases the first two stages");

__device__ __forceinline__ char* tm_pipe(char* pipe0) {                      // SWIZZLE_128B tiles need 1024-byte alignment
    return pipe0 + ((1024u - (smem_u32(pipe0) & 1023u)) & 1023u);
}
__device__ __forceinline__ void tm_init_barriers(char* pipe0) {              // once per kernel, one thread
    const unsigned bar = smem_u32(tm_pipe(pipe0) + kTmBarOff);
    for (int i = 0; i < 2 * kTmStages + 1; ++i) mbar_init(bar + 8u * i, 1);
}
__device__ __forceinline__ void tm_load_2d(unsigned dst, const void* tmap, int c0, int c1, unsigned mbar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(tmap), "r"(c0), "r"(c1), "r"(mbar) : "memory");
}
__device__ __forceinline__ unsigned long long tm_desc(unsigned smem_addr, unsigned lbo_bytes, unsigned sbo_bytes) {
    // cute::UMMA::SmemDescriptor: start[0,14) | LBO[16,30) | SBO[32,46) | version = 1 [46,48) | layout_type [61,64) = 2 (SWIZZLE_128B)
    return (unsigned long long)((smem_addr >> 4) & 0x3FFF) | ((unsigned long long)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((unsigned long long)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ void gemm_tile_tma(const Job& jb, int tile, char* pipe0, const RunArgs& ra, T5Ctx& cx) {
    char* pipe = tm_pipe(pipe0);
    const int M = (jb.rows_dynamic && ra.m_rows > 0) ? min(jb.M, ra.m_rows) : jb.M;
    const int N = jb.N, K = jb.K;
    const int tn = jb.tn;
    const int z = (tile >= jb.tpz) ? 1 : 0;
    const int rem = tile - z * jb.tpz;
    const int ti = rem / tn, tj = rem - ti * tn;
    const int i0 = ti * kBfM, j0 = tj * kBfN;
    if (i0 >= M) return;                                   // uniform per CTA
    const int i_valid = min(kBfM, M - i0), j_valid = min(kBfN, N - j0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool a_rc = jb.mode != G_DW, b_rc = jb.mode == G_FWD;
    float* s_mask = reinterpret_cast<float*>(pipe + kTmMaskOff);
    float* s_bias = reinterpret_cast<float*>(pipe + kTmEpiOff);
    float* s_w2 = s_bias + kBfN;
    const unsigned bar = smem_u32(pipe + kTmBarOff);       // full[s] = bar + 8s, empty[s] = bar + 8(S + s), done = bar + 16S
    const bool has_bias = jb.epi == E_BIAS || jb.epi == E_BIAS_RELU || jb.epi == E_BIAS_RELU_DOT;
    const int nch = (K + kBfK - 1) / kBfK;
    const unsigned it0 = cx.tm_it;

    if (warp == 0) {
        // ---- TMA producer: one thread, the whole K extent of this tile, up to kTmStages chunks ahead of the tensor core ----
        if (lane == 0) {
            const char* tmaps = reinterpret_cast<const char*>(ra.sh.tmaps);
            const void* tmA = tmaps + (size_t)(jb.tmA - 1 + z) * 128;
            const void* tmB = tmaps + (size_t)(jb.tmB - 1 + z) * 128;
            // operands were written through the generic proxy (other CTAs, previous phase; ordered by the grid barrier):
            // make them visible to the async proxy before the first bulk tensor read
            asm volatile("fence.proxy.async.global;" ::: "memory");
            for (int c = 0; c < nch; ++c) {
                const unsigned it = it0 + (unsigned)c, s = it % kTmStages, par = (it / kTmStages) & 1u;
                mbar_wait(bar + 8u * (kTmStages + s), par ^ 1u);          // the MMAs that read this stage last are done
                const unsigned full = bar + 8u * s;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full), "r"((unsigned)kTmStageBytes) : "memory");
                const unsigned dA = smem_u32(pipe + s * kTmStageBytes), dB = dA + kTmABytes;
                const int k0 = c * kBfK;
                if (a_rc) tm_load_2d(dA, tmA, k0, i0, full);
                else { tm_load_2d(dA, tmA, i0, k0, full); tm_load_2d(dA + 8192u, tmA, i0 + 64, k0, full); }
                if (b_rc) tm_load_2d(dB, tmB, k0, j0, full);
                else tm_load_2d(dB, tmB, j0, k0, full);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ---- MMA issuer: one thread ----
        if (lane == 0) {
            const unsigned idesc = bf_idesc() | ((a_rc ? 0u : 1u) << 15) | ((b_rc ? 0u : 1u) << 16);
            const unsigned a_step = a_rc ? 2u : 128u, b_step = b_rc ? 2u : 128u;      // 16-byte units per K=16 step
            for (int c = 0; c < nch; ++c) {
                const unsigned it = it0 + (unsigned)c, s = it % kTmStages, par = (it / kTmStages) & 1u;
                mbar_wait(bar + 8u * s, par);                                          // the chunk's boxes have landed
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const unsigned aA = smem_u32(pipe + s * kTmStageBytes);
                unsigned long long da = tm_desc(aA, a_rc ? 16u : 8192u, 1024u), db = tm_desc(aA + kTmABytes, b_rc ? 16u : 8192u, 1024u);
                const int kv = min(kBfK, K - c * kBfK), nk16 = (kv + 15) >> 4;
                for (int k16 = 0; k16 < nk16; ++k16) {
                    bf_mma(cx.tmem, da, db, idesc, (c | k16) ? 1u : 0u);
                    da += a_step; db += b_step;
                }
                t5_commit(bar + 8u * (kTmStages + s));                                 // stage free once these MMAs have read it
            }
            t5_commit(bar + 16u * kTmStages);                                          // tile done
        }
        __syncwarp();
    } else {
        // ---- the other warps: epilogue operands (their L2 latency hides behind the K loop) ----
        const int t = threadIdx.x - 64;                                                // 0..191
        if (t < kBfN) s_bias[t] = (has_bias && j0 + t < N) ? __ldcg(jb.bias + (size_t)z * jb.sBias + j0 + t) : 0.f;
        else if (t < 2 * kBfN) {
            const int c = t - kBfN;
            s_w2[c] = (jb.epi == E_BIAS_RELU_DOT && j0 + c < N) ? __ldcg(jb.aux0 + (size_t)z * jb.sAux + j0 + c) : 0.f;
        }
    }
    // relu mask of the tile (E_MASK): cp.async by the six idle warps
    const bool mask_vec = jb.epi == E_MASK && j_valid == kBfN && (jb.ldmask & 3) == 0 && (jb.sMask & 3) == 0;   // tile-uniform
    if (mask_vec && warp >= 2) {
        const float* mp = jb.mask + (size_t)z * jb.sMask + (size_t)i0 * jb.ldmask + j0;
        for (int v = threadIdx.x - 64; v < kBfM * 16; v += kThreads - 64) {
            const int o = v >> 4, r = (v & 15) << 2;
            cp_async16(s_mask + o * kBfPitchRC + r, o < i_valid ? mp + (size_t)o * jb.ldmask + r : mp, o < i_valid ? 16 : 0);
        }
    }
    cp_async_commit();
    cx.tm_it = it0 + (unsigned)nch;
    mbar_wait(bar + 16u * kTmStages, cx.tm_done);
    cx.tm_done ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    cp_async_wait<0>();
    __syncthreads();
    TR(206);

    // ---- epilogue (as gemm_tile_bf16): TMEM -> shared memory, then 16 threads per row: bias / relu / mask / Q-dot, fp32
    // store + the bf16 shadow of the output when a later GEMM reads it ----
    float* s_out = reinterpret_cast<float*>(pipe);             // [128][68]: stages 0-1 (every MMA of the tile has completed)
    {
        const int q4 = warp & 3, half = warp >> 2;
        unsigned r0[16], r1[16];
        const unsigned taddr = cx.tmem + ((unsigned)(32 * q4) << 16) + (unsigned)(half * 32);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                     : "=r"(r0[0]), "=r"(r0[1]), "=r"(r0[2]), "=r"(r0[3]), "=r"(r0[4]), "=r"(r0[5]), "=r"(r0[6]), "=r"(r0[7]), "=r"(r0[8]),
                       "=r"(r0[9]), "=r"(r0[10]), "=r"(r0[11]), "=r"(r0[12]), "=r"(r0[13]), "=r"(r0[14]), "=r"(r0[15])
                     : "r"(taddr));
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                     : "=r"(r1[0]), "=r"(r1[1]), "=r"(r1[2]), "=r"(r1[3]), "=r"(r1[4]), "=r"(r1[5]), "=r"(r1[6]), "=r"(r1[7]), "=r"(r1[8]),
                       "=r"(r1[9]), "=r"(r1[10]), "=r"(r1[11]), "=r"(r1[12]), "=r"(r1[13]), "=r"(r1[14]), "=r"(r1[15])
                     : "r"(taddr + 16u));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float* dst = s_out + (32 * q4 + lane) * kBfPitchRC + half * 32;
#pragma unroll
        for (int c = 0; c < 16; c += 4) {
            *reinterpret_cast<uint4*>(dst + c) = make_uint4(r0[c], r0[c + 1], r0[c + 2], r0[c + 3]);
            *reinterpret_cast<uint4*>(dst + 16 + c) = make_uint4(r1[c], r1[c + 1], r1[c + 2], r1[c + 3]);
        }
    }
    __syncthreads();
    {
        const bool relu = jb.epi == E_BIAS_RELU || jb.epi == E_BIAS_RELU_DOT;
        float* __restrict__ cbase = jb.C + (size_t)z * jb.sC + (size_t)i0 * jb.ldc + j0;
        unsigned short* c16 = jb.c16 ? ra.sh.base + (jb.c16 - 1) + (size_t)z * jb.sC16 + (size_t)i0 * jb.ldc16 + j0 : nullptr;
        const bool store_vec = j_valid == kBfN && (jb.ldc & 3) == 0 && ((reinterpret_cast<uintptr_t>(cbase) & 15) == 0);
        const int c4 = (threadIdx.x & 15) << 2;
        const float4 b4 = *reinterpret_cast<const float4*>(s_bias + c4), w4 = *reinterpret_cast<const float4*>(s_w2 + c4);
#pragma unroll
        for (int q = 0; q < (kBfM * 16) / kThreads; ++q) {       // 16 threads per row: 256 contiguous bytes
            const int row = (threadIdx.x >> 4) + 16 * q;
            const bool rok = row < i_valid;
            float4 x = *reinterpret_cast<const float4*>(s_out + row * kBfPitchRC + c4);
            x.x += b4.x; x.y += b4.y; x.z += b4.z; x.w += b4.w;
            if (relu) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
            if (jb.epi == E_MASK) {
                float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
                if (mask_vec) m = *reinterpret_cast<const float4*>(s_mask + row * kBfPitchRC + c4);
                else if (rok) {
                    const float* mp = jb.mask + (size_t)z * jb.sMask + (size_t)(i0 + row) * jb.ldmask + j0 + c4;
                    if (c4 + 0 < j_valid) m.x = __ldcg(mp); if (c4 + 1 < j_valid) m.y = __ldcg(mp + 1);
                    if (c4 + 2 < j_valid) m.z = __ldcg(mp + 2); if (c4 + 3 < j_valid) m.w = __ldcg(mp + 3);
                }
                x.x = m.x > 0.f ? x.x : 0.f; x.y = m.y > 0.f ? x.y : 0.f; x.z = m.z > 0.f ? x.z : 0.f; x.w = m.w > 0.f ? x.w : 0.f;
            }
            if (c4 + 0 >= j_valid) x.x = 0.f;
            if (c4 + 1 >= j_valid) x.y = 0.f;
            if (c4 + 2 >= j_valid) x.z = 0.f;
            if (c4 + 3 >= j_valid) x.w = 0.f;
            if (jb.epi == E_BIAS_RELU_DOT) {                     // Q-dot partial of this row over the tile's 64 columns
                float dot = fmaf(x.x, w4.x, fmaf(x.y, w4.y, fmaf(x.z, w4.z, x.w * w4.w)));
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
                if (rok && (threadIdx.x & 15) == 0) jb.out0[((size_t)z * tn + tj) * jb.i0 + i0 + row] = dot;
            }
            if (rok) {
                float* dst = cbase + (size_t)row * jb.ldc + c4;
                if (store_vec) *reinterpret_cast<float4*>(dst) = x;
                else {
                    if (c4 + 0 < j_valid) dst[0] = x.x;
                    if (c4 + 1 < j_valid) dst[1] = x.y;
                    if (c4 + 2 < j_valid) dst[2] = x.z;
                    if (c4 + 3 < j_valid) dst[3] = x.w;
                }
                if (c16) {
                    unsigned short* d16 = c16 + (size_t)row * jb.ldc16 + c4;
                    if (c4 + 3 < j_valid) *reinterpret_cast<uint2*>(d16) = make_uint2(bf16x2_rn(x.x, x.y), bf16x2_rn(x.z, x.w));
                    else {
                        if (c4 + 0 < j_valid) d16[0] = bf16_rn(x.x);
                        if (c4 + 1 < j_valid) d16[1] = bf16_rn(x.y);
                        if (c4 + 2 < j_valid) d16[2] = bf16_rn(x.z);
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();                           // TMEM and smem are reused by the next tile
    TR(207);
}
