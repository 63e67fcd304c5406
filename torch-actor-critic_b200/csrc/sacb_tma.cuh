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
constexpr int kTmAcc = 4;                  // independent TMEM accumulators (k16 step mod 4), summed in the epilogue
constexpr int kTmABytes = kBfM * kBfK * 2, kTmBBytes = kBfN * kBfK * 2;      // 16,384 + 8,192
constexpr int kTmStageBytes = kTmABytes + kTmBBytes;
constexpr int kTmMaskOff = kTmStages * kTmStageBytes;                        // 147,456: relu-mask tile [128][68] fp32
constexpr int kTmEpiOff = kTmMaskOff + kBfM * kBfPitchRC * 4;                // 182,272: bias[64], w2[64]
constexpr int kTmBarOff = kTmEpiOff + 2 * kBfN * 4;                          // 182,784: full[6], empty[6], done
constexpr int kTmMaskPitch16 = 72;                                           // bf16 mask tile pitch (144 B rows: 16-byte aligned, conflict-free)
constexpr int kTmEnd = kTmBarOff + (2 * kTmStages + 1) * 8;
static_assert(kTmEnd + 1024 <= kT5PipeBytes, "TMA tile buffers (+ 1 KB alignment slack) must fit in the tcgen05 kernel's shared memory");
static_assert(kBfM * kBfPitchRC * 4 <= 2 * kTmStageBytes, "the epilogue's tile buffer aliases the first two stages");

__device__ __forceinline__ char* tm_pipe(char* pipe0) {                      // SWIZZLE_128B tiles need 1024-byte alignment
    return pipe0 + ((1024u - (smem_u32(pipe0) & 1023u)) & 1023u);
}
__device__ __forceinline__ void tm_init_barriers(char* pipe0) {              // once per kernel, one thread
    const unsigned bar = smem_u32(tm_pipe(pipe0) + kTmBarOff);
    for (int i = 0; i < 2 * kTmStages + 1; ++i) mbar_init(bar + 8u * i, i < kTmStages ? 2 : 1);   // full[s]: one arrival per operand
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

// Epilogue of a full tile (64 valid columns, aligned outputs), specialised per epilogue kind: every job field it needs is in
// a register and the eight rows of a thread are processed together -- 8 shared-memory loads, then the math, then the
// stores -- instead of eight trips through a chain of branches on fields re-read from the job table (measured: 3,500
// cycles per tile for the generic loop below, of a 13,000-cycle tile).
template <int EPI>
__device__ __forceinline__ void tm_epilogue_full(const float* s_out, const float* s_mask, const unsigned short* s_mask16, bool m16,
                                                 float4 b4, float4 w4, float* cbase, int ldc, unsigned short* c16, int ldc16,
                                                 float* dotp, int i_valid, int j_valid, bool vec32) {
    const int r0 = threadIdx.x >> 4, c4 = (threadIdx.x & 15) << 2;
    constexpr int Q = (kBfM * 16) / kThreads;                       // 8 rows per thread
    float4 x[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) x[q] = *reinterpret_cast<const float4*>(s_out + (r0 + 16 * q) * kBfPitchRC + c4);
    if (EPI == E_MASK) {
        if (m16) {
            uint2 mb[Q];
#pragma unroll
            for (int q = 0; q < Q; ++q) mb[q] = *reinterpret_cast<const uint2*>(s_mask16 + (r0 + 16 * q) * kTmMaskPitch16 + c4);
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                x[q].x = (short)(mb[q].x & 0xFFFFu) > 0 ? x[q].x : 0.f; x[q].y = (short)(mb[q].x >> 16) > 0 ? x[q].y : 0.f;
                x[q].z = (short)(mb[q].y & 0xFFFFu) > 0 ? x[q].z : 0.f; x[q].w = (short)(mb[q].y >> 16) > 0 ? x[q].w : 0.f;
            }
        } else {
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                const float4 m = *reinterpret_cast<const float4*>(s_mask + (r0 + 16 * q) * kBfPitchRC + c4);
                x[q].x = m.x > 0.f ? x[q].x : 0.f; x[q].y = m.y > 0.f ? x[q].y : 0.f;
                x[q].z = m.z > 0.f ? x[q].z : 0.f; x[q].w = m.w > 0.f ? x[q].w : 0.f;
            }
        }
    }
    if (EPI == E_BIAS || EPI == E_BIAS_RELU || EPI == E_BIAS_RELU_DOT) {
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            x[q].x += b4.x; x[q].y += b4.y; x[q].z += b4.z; x[q].w += b4.w;
            if (EPI != E_BIAS) { x[q].x = fmaxf(x[q].x, 0.f); x[q].y = fmaxf(x[q].y, 0.f); x[q].z = fmaxf(x[q].z, 0.f); x[q].w = fmaxf(x[q].w, 0.f); }
        }
    }
    if (EPI == E_BIAS_RELU_DOT) {                                  // Q-dot partials of the 8 rows over the tile's 64 columns
        float dot[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) dot[q] = fmaf(x[q].x, w4.x, fmaf(x[q].y, w4.y, fmaf(x[q].z, w4.z, x[q].w * w4.w)));
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
#pragma unroll
            for (int q = 0; q < Q; ++q) dot[q] += __shfl_xor_sync(0xffffffffu, dot[q], o);
        }
        if ((threadIdx.x & 15) == 0) {
#pragma unroll
            for (int q = 0; q < Q; ++q) if (r0 + 16 * q < i_valid) dotp[r0 + 16 * q] = dot[q];
        }
    }
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        const int row = r0 + 16 * q;
        if (row < i_valid) {
            if (cbase) {
                float* dst = cbase + (size_t)row * ldc + c4;
                if (vec32) *reinterpret_cast<float4*>(dst) = x[q];
                else {                                   // ragged / unaligned output (e.g. a [.., 393] gradient): element stores
                    if (c4 + 0 < j_valid) dst[0] = x[q].x;
                    if (c4 + 1 < j_valid) dst[1] = x[q].y;
                    if (c4 + 2 < j_valid) dst[2] = x[q].z;
                    if (c4 + 3 < j_valid) dst[3] = x[q].w;
                }
            }
            if (c16) {
                unsigned short* d16 = c16 + (size_t)row * ldc16 + c4;
                if (c4 + 3 < j_valid) *reinterpret_cast<uint2*>(d16) = make_uint2(bf16x2_rn(x[q].x, x[q].y), bf16x2_rn(x[q].z, x[q].w));
                else {
                    if (c4 + 0 < j_valid) d16[0] = bf16_rn(x[q].x);
                    if (c4 + 1 < j_valid) d16[1] = bf16_rn(x[q].y);
                    if (c4 + 2 < j_valid) d16[2] = bf16_rn(x[q].z);
                }
            }
        }
    }
}

// Issued by thread 0 between the grid barrier's arrive and its wait: operand B of this CTA's first tile of the NEXT phase,
// when the previous phase did not write it (Job.b_static: weights, or forward activations during the backward pass).  The
// first kTmStages chunks of B (and the tensor-map fetch) then overlap the barrier; the tile's producer skips them.
__device__ __forceinline__ void tm_prefetch_b(const Job& jb, int tile, char* pipe0, const RunArgs& ra, T5Ctx& cx) {
    char* pipe = tm_pipe(pipe0);
    const int z = (tile >= jb.tpz) ? 1 : 0;
    const int rem = tile - z * jb.tpz;
    const int ti = rem / jb.tn, tj = rem - ti * jb.tn;
    if (ti * kBfM >= jb.M) return;
    const int j0 = tj * kBfN;
    const bool b_rc = jb.mode == G_FWD;
    const unsigned bar = smem_u32(pipe + kTmBarOff);
    const char* tmaps = reinterpret_cast<const char*>(ra.sh.tmaps);
    const void* tmB = tmaps + (size_t)(jb.tmB - 1 + z) * 128;
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmaps + (size_t)(jb.tmA - 1 + z) * 128) : "memory");
    const int nch = (jb.K + kBfK - 1) / kBfK, npre = nch < kTmStages ? nch : kTmStages;
    for (int c = 0; c < npre; ++c) {
        const unsigned it = cx.tm_it + (unsigned)c, s = it % kTmStages, par = (it / kTmStages) & 1u;
        mbar_wait(bar + 8u * (kTmStages + s), par ^ 1u);
        const unsigned full = bar + 8u * s;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full), "r"((unsigned)kTmBBytes) : "memory");
        const unsigned dB = smem_u32(pipe + s * kTmStageBytes) + kTmABytes;
        if (b_rc) tm_load_2d(dB, tmB, c * kBfK, j0, full);
        else tm_load_2d(dB, tmB, j0, c * kBfK, full);
    }
}

__device__ __forceinline__ void gemm_tile_tma(const Job& jb, int tile, char* pipe0, const RunArgs& ra, T5Ctx& cx, bool b_pre) {
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
    TR(200);

    if (warp == 0) {
        // ---- TMA producer: one thread, the whole K extent of this tile, up to kTmStages chunks ahead of the tensor core ----
        if (lane == 0) {
            const char* tmaps = reinterpret_cast<const char*>(ra.sh.tmaps);
            const void* tmA = tmaps + (size_t)(jb.tmA - 1 + z) * 128;
            const void* tmB = tmaps + (size_t)(jb.tmB - 1 + z) * 128;
            // operands were written through the generic proxy (other CTAs, previous phase; ordered by the grid barrier):
            // make them visible to the async proxy before the first bulk tensor read
            asm volatile("fence.proxy.async.global;" ::: "memory");
            TR(201);
            const int pre = b_pre ? (nch < kTmStages ? nch : kTmStages) : 0;     // chunks whose B operand tm_prefetch_b already requested
            for (int c = 0; c < nch; ++c) {
                const unsigned it = it0 + (unsigned)c, s = it % kTmStages, par = (it / kTmStages) & 1u;
                mbar_wait(bar + 8u * (kTmStages + s), par ^ 1u);          // the MMAs that read this stage last are done
                const unsigned full = bar + 8u * s;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full), "r"((unsigned)kTmABytes) : "memory");
                const unsigned dA = smem_u32(pipe + s * kTmStageBytes), dB = dA + kTmABytes;
                const int k0 = c * kBfK;
                if (a_rc) tm_load_2d(dA, tmA, k0, i0, full);
                else { tm_load_2d(dA, tmA, i0, k0, full); tm_load_2d(dA + 8192u, tmA, i0 + 64, k0, full); }
                if (c >= pre) {
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full), "r"((unsigned)kTmBBytes) : "memory");
                    if (b_rc) tm_load_2d(dB, tmB, k0, j0, full);
                    else tm_load_2d(dB, tmB, j0, k0, full);
                }
                if (c == 0) TR(202);
#ifdef SACB_TRACE_FIRST
                if (c == 0) { mbar_wait(full, par); TR(210); }          // experiment: time from first issue to first landing
#endif
            }
            TR(203);
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
                // Four independent accumulators, one per k16 step of a chunk: an MMA into the accumulator the previous one just
                // wrote waits for that write (~130 cycles each for a 128x64x16 MMA whose math is ~32 cycles -- the K loop
                // of a K=512 tile measured ~4,500 cycles on one accumulator, as long as the whole operand stream).
                for (int k16 = 0; k16 < nk16; ++k16) {
                    bf_mma(cx.tmem + (unsigned)(k16 * kBfN), da, db, idesc, c ? 1u : 0u);
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
    // relu mask of the tile (E_MASK): cp.async by the six idle warps.  From the bf16 shadow of the activation when there is
    // one (only its sign is needed; [128][72] bf16), else the fp32 tile
    const bool mask_vec = jb.epi == E_MASK && !jb.mask16 && j_valid == kBfN && (jb.ldmask & 3) == 0 && (jb.sMask & 3) == 0;   // tile-uniform
    unsigned short* s_mask16 = reinterpret_cast<unsigned short*>(s_mask);
    if (jb.epi == E_MASK && jb.mask16 && warp >= 2) {
        const unsigned short* mp = ra.sh.base + (jb.mask16 - 1) + (size_t)z * jb.sMask + (size_t)i0 * jb.ldmask16 + j0;
        for (int v = threadIdx.x - 64; v < kBfM * 8; v += kThreads - 64) {
            const int o = v >> 3, r = (v & 7) << 3;
            const bool in = o < i_valid && j0 + r < jb.ldmask16;          // the shadow's pitch is padded to 8 elements (zeros)
            cp_async16(s_mask16 + o * kTmMaskPitch16 + r, in ? mp + (size_t)o * jb.ldmask16 + r : mp, in ? 16 : 0);
        }
    }
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
    TR(204);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    cp_async_wait<0>();
    __syncthreads();
    TR(206);

    // ---- epilogue (as gemm_tile_bf16): TMEM -> shared memory, then 16 threads per row: bias / relu / mask / Q-dot, fp32
    // store + the bf16 shadow of the output when a later GEMM reads it ----
    float* s_out = reinterpret_cast<float*>(pipe);             // [128][68]: stages 0-1 (every MMA of the tile has completed)
    {
        const int q4 = warp & 3, half = warp >> 2;
        const int nacc = min(kTmAcc, (min(K, kBfK) + 15) >> 4);          // accumulators the first chunk initialised
        float v0[16], v1[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) { v0[c] = 0.f; v1[c] = 0.f; }
        for (int a = 0; a < nacc; ++a) {                                 // fixed summation order -> deterministic
            unsigned r0[16], r1[16];
            const unsigned taddr = cx.tmem + ((unsigned)(32 * q4) << 16) + (unsigned)(a * kBfN + half * 32);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                         : "=r"(r0[0]), "=r"(r0[1]), "=r"(r0[2]), "=r"(r0[3]), "=r"(r0[4]), "=r"(r0[5]), "=r"(r0[6]), "=r"(r0[7]), "=r"(r0[8]),
                           "=r"(r0[9]), "=r"(r0[10]), "=r"(r0[11]), "=r"(r0[12]), "=r"(r0[13]), "=r"(r0[14]), "=r"(r0[15])
                         : "r"(taddr));
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                         : "=r"(r1[0]), "=r"(r1[1]), "=r"(r1[2]), "=r"(r1[3]), "=r"(r1[4]), "=r"(r1[5]), "=r"(r1[6]), "=r"(r1[7]), "=r"(r1[8]),
                           "=r"(r1[9]), "=r"(r1[10]), "=r"(r1[11]), "=r"(r1[12]), "=r"(r1[13]), "=r"(r1[14]), "=r"(r1[15])
                         : "r"(taddr + 16u));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int c = 0; c < 16; ++c) { v0[c] += __uint_as_float(r0[c]); v1[c] += __uint_as_float(r1[c]); }
        }
        float* dst = s_out + (32 * q4 + lane) * kBfPitchRC + half * 32;
#pragma unroll
        for (int c = 0; c < 16; c += 4) {
            *reinterpret_cast<float4*>(dst + c) = make_float4(v0[c], v0[c + 1], v0[c + 2], v0[c + 3]);
            *reinterpret_cast<float4*>(dst + 16 + c) = make_float4(v1[c], v1[c + 1], v1[c + 2], v1[c + 3]);
        }
    }
    __syncthreads();
    TR(208);
    float* __restrict__ cbase = jb.C + (size_t)z * jb.sC + (size_t)i0 * jb.ldc + j0;
    unsigned short* c16 = jb.c16 ? ra.sh.base + (jb.c16 - 1) + (size_t)z * jb.sC16 + (size_t)i0 * jb.ldc16 + j0 : nullptr;
    const bool store_vec = j_valid == kBfN && (jb.ldc & 3) == 0 && ((reinterpret_cast<uintptr_t>(cbase) & 15) == 0);
    const int c4 = (threadIdx.x & 15) << 2;
    const float4 b4 = *reinterpret_cast<const float4*>(s_bias + c4), w4 = *reinterpret_cast<const float4*>(s_w2 + c4);
    const int epi = jb.epi;
    const bool m16 = jb.mask16 != 0, nc32 = jb.nc32 != 0;
    if (epi != E_MASK || m16 || mask_vec) {        // (a ragged fp32 mask tile takes the generic loop below)
        float* cb = nc32 ? nullptr : cbase;
        float* dotp = epi == E_BIAS_RELU_DOT ? jb.out0 + ((size_t)z * tn + tj) * jb.i0 + i0 : nullptr;
        const int ldc = jb.ldc, ldc16 = jb.ldc16;
        switch (epi) {
            case E_STORE: tm_epilogue_full<E_STORE>(s_out, s_mask, s_mask16, m16, b4, w4, cb, ldc, c16, ldc16, dotp, i_valid, j_valid, store_vec); break;
            case E_BIAS: tm_epilogue_full<E_BIAS>(s_out, s_mask, s_mask16, m16, b4, w4, cb, ldc, c16, ldc16, dotp, i_valid, j_valid, store_vec); break;
            case E_BIAS_RELU: tm_epilogue_full<E_BIAS_RELU>(s_out, s_mask, s_mask16, m16, b4, w4, cb, ldc, c16, ldc16, dotp, i_valid, j_valid, store_vec); break;
            case E_BIAS_RELU_DOT: tm_epilogue_full<E_BIAS_RELU_DOT>(s_out, s_mask, s_mask16, m16, b4, w4, cb, ldc, c16, ldc16, dotp, i_valid, j_valid, store_vec); break;
            default: tm_epilogue_full<E_MASK>(s_out, s_mask, s_mask16, m16, b4, w4, cb, ldc, c16, ldc16, dotp, i_valid, j_valid, store_vec); break;
        }
    } else {
        const bool relu = jb.epi == E_BIAS_RELU || jb.epi == E_BIAS_RELU_DOT;
#pragma unroll
        for (int q = 0; q < (kBfM * 16) / kThreads; ++q) {       // 16 threads per row: 256 contiguous bytes
            const int row = (threadIdx.x >> 4) + 16 * q;
            const bool rok = row < i_valid;
            float4 x = *reinterpret_cast<const float4*>(s_out + row * kBfPitchRC + c4);
            x.x += b4.x; x.y += b4.y; x.z += b4.z; x.w += b4.w;
            if (relu) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f); }
            if (jb.epi == E_MASK) {
                float4 m = make_float4(0.f, 0.f, 0.f, 0.f);
                if (jb.mask16) {                                 // signs of the bf16 activation: > 0 <=> positive as int16
                    const uint2 mb = *reinterpret_cast<const uint2*>(s_mask16 + row * kTmMaskPitch16 + c4);
                    m.x = (short)(mb.x & 0xFFFFu) > 0 ? 1.f : 0.f; m.y = (short)(mb.x >> 16) > 0 ? 1.f : 0.f;
                    m.z = (short)(mb.y & 0xFFFFu) > 0 ? 1.f : 0.f; m.w = (short)(mb.y >> 16) > 0 ? 1.f : 0.f;
                } else if (mask_vec) m = *reinterpret_cast<const float4*>(s_mask + row * kBfPitchRC + c4);
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
                if (jb.nc32) {}
                else if (store_vec) *reinterpret_cast<float4*>(dst) = x;
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
    TR(209);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();                           // TMEM and smem are reused by the next tile
    TR(207);
}
