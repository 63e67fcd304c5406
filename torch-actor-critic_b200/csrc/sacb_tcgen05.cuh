// sacb_tcgen05.cuh -- Blackwell-native GEMM tile: tcgen05.mma (5th-gen tensor cores) with the accumulator in TMEM.
// Included by sacb_device.cuh (inside namespace sacb).
//
//   C[64 x 16] (+epilogue) = sum_r A(i,r) * B(j,r), fp32 storage, 3xTF32 error compensation:
//   every operand tile is split in shared memory into hi = tf32(x) and lo = tf32(x - hi) and the tensor core
//   accumulates a_lo*b_hi + a_hi*b_lo + a_hi*b_hi in fp32 (TMEM) -> fp32-level accuracy.
//
// Operand tiles live in shared memory in canonical UMMA layouts (cute/atom/mma_traits_sm100.hpp, make_umma_desc),
// built from 16-byte chunks (4 tf32) so that a 16-byte cp.async drops every global chunk straight into place:
//   K-major  (reduction contiguous in global memory), SWIZZLE_NONE:
//       byte(m, k) = (m/8)*SBO + (k/4)*LBO + (m%8)*16 + (k%4)*4            LBO = 128, SBO = 8192
//       one tcgen05.mma (K = 8 tf32) reads two chunk columns: start += 256 B per k-step
//   MN-major (output contiguous in global memory): tf32 supports ONLY SWIZZLE_128B_BASE32B here
//   (cutlass/gemm/collective/builders/sm100_common.inl:92).  Atom = 4 k-rows x 32 outputs (128 B per k-row), the four
//   32-byte groups of a row XOR-permuted by k%4 (Swizzle<2,5,2>):
//       byte(m, k) = (m/32)*LBO + (k/4)*SBO + (k%4)*128 + 32*(((m%32)/8) ^ (k%4)) + (m%8)*4
//       LBO = 512 (next 32-output atom), SBO = 512 * atoms (next group of 4 k);  start += 2*SBO per k-step
// Accumulator: M = 64 rows occupy TMEM lanes (m%16) + 32*(m/16) (half of each sub-partition), 16 fp32 columns;
// warp w reads rows 16w..16w+15 with tcgen05.ld.32x32b.x16 (lanes 0..15 hold one row each).
#pragma once

constexpr int kT5M = 64, kT5N = 16, kT5K = 256;                  // tile and K block
constexpr int kT5Acc = 8, kT5TmemCols = kT5Acc * kT5N;           // independent TMEM accumulators (128 columns)
constexpr int kT5AHi = 0, kT5ALo = 65536, kT5BHi = 131072, kT5BLo = 163840, kT5Epi = 196608, kT5PipeBytes = 198656;
constexpr int kT5SboK = 8192, kT5LboK = 128;                     // K-major strides (bytes)
constexpr int kT5LboMN = 512;                                    // MN-major: 32-output atom stride; SBO = 512 * atoms

__device__ __forceinline__ unsigned long long t5_desc(unsigned smem_addr, unsigned lbo_bytes, unsigned sbo_bytes, bool mn_major) {
    // cute::UMMA::SmemDescriptor: start[0,14) | LBO[16,30) | SBO[32,46) | version=1 [46,48) | layout_type [61,64):
    // SWIZZLE_NONE = 0 (K-major here), SWIZZLE_128B_BASE32B = 1 (MN-major tf32)
    return (unsigned long long)((smem_addr >> 4) & 0x3FFF) | ((unsigned long long)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((unsigned long long)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | ((mn_major ? 1ull : 0ull) << 61);
}
__device__ __forceinline__ unsigned t5_idesc(bool a_mn, bool b_mn) {
    // cute::UMMA::InstrDescriptor: c_format F32 (1) [4,6) | a_format TF32 (2) [7,10) | b_format TF32 (2) [10,13) |
    // a_major [15] | b_major [16] | N>>3 [17,23) | M>>4 [24,29)
    return (1u << 4) | (2u << 7) | (2u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((unsigned)(kT5N >> 3) << 17) | ((unsigned)(kT5M >> 4) << 24);
}
__device__ __forceinline__ void t5_mma(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned idesc, unsigned acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void t5_commit(unsigned mbar_addr) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar_addr) : "memory");
}
__device__ __forceinline__ void mbar_init(unsigned addr, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned addr, unsigned parity) {
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n" ::"r"(addr), "r"(parity) : "memory");
}

// byte offset of element (o, k) of an operand tile with R outputs
template <bool RC>
__device__ __forceinline__ int t5_off(int o, int k, int R) {
    if (RC) return (o >> 3) * kT5SboK + (k >> 2) * kT5LboK + (o & 7) * 16 + (k & 3) * 4;
    const int sbo = ((R + 31) >> 5) * 512;
    return (o >> 5) * kT5LboMN + (k >> 2) * sbo + (k & 3) * 128 + 32 * (((o & 31) >> 3) ^ (k & 3)) + (o & 7) * 4;
}

// chunk loader.  RC = K-major.  R = tile outputs (64 for A, 16 for B), compile-time so that all index math is shifts.
// Work unit = 8 rows (K-major) / 8 k's (MN-major) x 4 chunks = one warp instruction: the 8 lanes that share a chunk
// column write 8 consecutive 16-byte slots (conflict-free), the 4 chunk columns read 64 contiguous bytes per row.
template <bool RC, int R>
__device__ __forceinline__ void t5_load(char* buf, const float* __restrict__ g, int ld, int o0, int o_valid,
                                        int k0, int k_valid, int vec) {
    const int k_pad = (k_valid + 7) & ~7;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int l8 = lane & 7, c4 = lane >> 3;
    if (vec) {
        if (RC) {
            constexpr int MG = R >> 3;                     // 8 (A) or 2 (B) row groups
            const int CB = (k_pad + 15) >> 4;              // chunk blocks of 4 chunks = 16 k
            const int mg = warp & (MG - 1);
            const int row = mg * 8 + l8;
            const bool row_ok = row < o_valid;
            const float* src_row = g + (size_t)(o0 + row) * ld + k0;
            char* dst_row = buf + mg * kT5SboK + l8 * 16;
#pragma unroll 4
            for (int cb = warp / MG; cb < CB; cb += kWarps / MG) {
                const int kk = (cb * 4 + c4) * 4;
                if (kk < k_pad) {
                    int nb = 0;
                    if (row_ok && kk < k_valid) nb = min(4, k_valid - kk) * 4;
                    cp_async16(dst_row + (kk >> 2) * kT5LboK, nb ? src_row + kk : g, nb);
                }
            }
        } else {
            constexpr int CB = R >> 4;                     // 4 (A) or 1 (B) chunk blocks of 16 outputs
            constexpr int SBO = ((R + 31) >> 5) * 512;
            const int KG = k_pad >> 3;
            const int cbi = warp & (CB - 1);
            const int oo = (cbi * 4 + c4) * 4;
            int nb_o = 0;
            if (oo < o_valid) nb_o = min(4, o_valid - oo) * 4;
            const float* src_col = g + (size_t)k0 * ld + o0 + oo;
            // byte(o,k) = (o/32)*512 + (k/4)*SBO + (k%4)*128 + 32*(((o%32)/8) ^ (k%4)) + (o%8)*4
            char* dst_col = buf + (oo >> 5) * kT5LboMN + (oo & 7) * 4;
            const int og = (oo & 31) >> 3;
#pragma unroll 4
            for (int kg = warp / CB; kg < KG; kg += kWarps / CB) {
                const int kk = kg * 8 + l8;
                const int nb = kk < k_valid ? nb_o : 0;
                cp_async16(dst_col + (kk >> 2) * SBO + (kk & 3) * 128 + 32 * (og ^ (kk & 3)), nb ? src_col + (size_t)kk * ld : g, nb);
            }
        }
    } else {            // element-wise path for unaligned operands; outputs beyond o_valid stay stale (feed only unused outputs)
        const int n_el = o_valid * k_pad;
        for (int e = threadIdx.x; e < n_el; e += kThreads) {
            int o, k;
            if (RC) { o = e / k_pad; k = e - o * k_pad; } else { k = e / o_valid; o = e - k * o_valid; }
            char* dst = buf + t5_off<RC>(o, k, R);
            if (k < k_valid) cp_async4(dst, RC ? g + (size_t)(o0 + o) * ld + k0 + k : g + (size_t)(k0 + k) * ld + o0 + o);
            else *reinterpret_cast<float*>(dst) = 0.f;
        }
    }
}

__device__ __forceinline__ void t5_split4(char* hi, char* lo, int off) {
    float4 x = *reinterpret_cast<const float4*>(hi + off);
    // hi = x with the 13 low mantissa bits cleared (exactly a tf32 value, whatever rounding the tensor core applies);
    // lo = x - hi is exact in fp32 and |lo| < 2^-10 |x|, so its own tf32 truncation costs < 2^-20 |x|.
    float4 h, l;
    h.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u); l.x = x.x - h.x;
    h.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u); l.y = x.y - h.y;
    h.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u); l.z = x.z - h.z;
    h.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u); l.w = x.w - h.w;
    *reinterpret_cast<float4*>(hi + off) = h;
    *reinterpret_cast<float4*>(lo + off) = l;
}

// hi/lo split of the region the loader touched: x -> hi (in place), lo -> buf_lo (same layout).
template <bool RC, int R>
__device__ __forceinline__ void t5_split(char* hi, char* lo, int k_valid) {
    const int k_pad = (k_valid + 7) & ~7;
    if (RC) {           // per 8-row group: (k_pad/4) chunk columns x 128 B, contiguous
        const int bytes = (k_pad >> 2) * 128;
#pragma unroll
        for (int mg = 0; mg < (R >> 3); ++mg) {
#pragma unroll 2
            for (int off = threadIdx.x * 16; off < bytes; off += kThreads * 16) t5_split4(hi, lo, mg * kT5SboK + off);
        }
    } else {            // contiguous: (k_pad/4) k-groups x atoms x 512 B
        const int bytes = (k_pad >> 2) * ((R + 31) >> 5) * 512;
#pragma unroll 4
        for (int off = threadIdx.x * 16; off < bytes; off += kThreads * 16) t5_split4(hi, lo, off);
    }
}

struct T5Ctx {
    unsigned tmem; unsigned mbar; unsigned phase;     // TMEM base address, mbarrier smem address, its parity
    unsigned mbar_op0, mbar_op1, phase_op;            // bf16 path: one mbarrier per operand buffer; parity bit b of phase_op
    unsigned tm_it, tm_done;                               // TMA path (sacb_tma.cuh): K chunks issued so far (stage ring position), parity of `done`
};

template <bool A_RC, bool B_RC>
__device__ __forceinline__ void gemm_tile_t5(const Job& jb, int tile, char* pipe, const RunArgs& ra, T5Ctx& cx) {
    TR(100);
    const int M = (jb.rows_dynamic && ra.m_rows > 0) ? min(jb.M, ra.m_rows) : jb.M;
    const int N = jb.N, K = jb.K;
    const int tn = jb.tn;
    const int z = (tile >= jb.tpz) ? 1 : 0;
    const int rem = tile - z * jb.tpz;
    const int ti = rem / tn, tj = rem - ti * tn;
    const int i0 = ti * kT5M, j0 = tj * kT5N;
    if (i0 >= M) return;                                   // uniform per CTA
    const int i_valid = min(kT5M, M - i0), j_valid = min(kT5N, N - j0);
    const float* __restrict__ Ag = jb.A + (size_t)z * jb.sA;
    const float* __restrict__ Bg = jb.B + (size_t)z * jb.sB;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    char* a_hi = pipe + kT5AHi; char* a_lo = pipe + kT5ALo; char* b_hi = pipe + kT5BHi; char* b_lo = pipe + kT5BLo;
    float* s_bias = reinterpret_cast<float*>(pipe + kT5Epi);      // [16] bias, [16] w2, [16][17] post-product X slice
    float* s_w2 = s_bias + 16;
    float* s_x = s_w2 + 16;

    const unsigned idesc = t5_idesc(!A_RC, !B_RC);
    constexpr unsigned a_lbo = A_RC ? kT5LboK : kT5LboMN, a_sbo = A_RC ? kT5SboK : 512u * ((kT5M + 31) / 32);
    constexpr unsigned b_lbo = B_RC ? kT5LboK : kT5LboMN, b_sbo = B_RC ? kT5SboK : 512u * ((kT5N + 31) / 32);
    constexpr unsigned a_step = A_RC ? 256u : 2u * a_sbo, b_step = B_RC ? 256u : 2u * b_sbo;      // bytes per k-step of 8

    // this thread's accumulator row (warps 0..3, lanes 0..15: row = 16*warp + lane)
    const int erow = 16 * warp + lane;
    const bool e_thread = warp < 4 && lane < 16;
    const int gi = i0 + erow;
    const bool row_ok = e_thread && gi < M;
    const bool has_bias = jb.epi == E_BIAS || jb.epi == E_BIAS_RELU || jb.epi == E_BIAS_RELU_DOT;
    const int J = jb.i1;

    for (int kb = 0; kb < K; kb += kT5K) {
        const int kv = min(kT5K, K - kb);
        t5_load<A_RC, kT5M>(a_hi, Ag, jb.lda, i0, i_valid, kb, kv, jb.a_vec);
        t5_load<B_RC, kT5N>(b_hi, Bg, jb.ldb, j0, j_valid, kb, kv, jb.b_vec);
        cp_async_commit();
        if (kb == 0) {
            // epilogue operands of this tile's 16 columns -> shared memory (their L2 latency hides behind the operand loads)
            const int t = threadIdx.x;
            if (t < 16) s_bias[t] = (has_bias && j0 + t < N) ? __ldcg(jb.bias + (size_t)z * jb.sBias + j0 + t) : 0.f;
            else if (t < 32) s_w2[t - 16] = (jb.epi == E_BIAS_RELU_DOT && j0 + t - 16 < N) ? __ldcg(jb.aux0 + (size_t)z * jb.sAux + j0 + t - 16) : 0.f;
            if (J > 0 && t < 16 * J) {
                const int j = t >> 4, c = t & 15;
                const float* __restrict__ X = jb.aux1 + (size_t)z * jb.sAux;
                s_x[j * 17 + c] = (j0 + c < N) ? __ldcg(jb.i3 ? X + (size_t)(j0 + c) * jb.i2 + j : X + (size_t)j * jb.i2 + j0 + c) : 0.f;
            }
        }
        cp_async_wait<0>();
        __syncthreads();
        TR(110);
        t5_split<A_RC, kT5M>(a_hi, a_lo, kv);
        t5_split<B_RC, kT5N>(b_hi, b_lo, kv);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy smem writes -> visible to the tensor core
        __syncthreads();
        TR(111);
        if (threadIdx.x == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            unsigned long long dah = t5_desc(smem_u32(a_hi), a_lbo, a_sbo, !A_RC), dal = t5_desc(smem_u32(a_lo), a_lbo, a_sbo, !A_RC);
            unsigned long long dbh = t5_desc(smem_u32(b_hi), b_lbo, b_sbo, !B_RC), dbl = t5_desc(smem_u32(b_lo), b_lbo, b_sbo, !B_RC);
            const int nk8 = (kv + 7) >> 3;
            // kT5Acc independent accumulators (k-step mod kT5Acc): MMAs into the same TMEM tile serialise on its ~64-cycle
            // latency, and with N = 16 each MMA is only ~8 cycles of math
#pragma unroll 4
            for (int k8 = 0; k8 < nk8; ++k8) {
                const unsigned d = cx.tmem + (unsigned)((k8 & (kT5Acc - 1)) * kT5N);
                t5_mma(d, dal, dbh, idesc, (kb == 0 && k8 < kT5Acc) ? 0u : 1u);
                t5_mma(d, dah, dbl, idesc, 1u);
                t5_mma(d, dah, dbh, idesc, 1u);
                dah += a_step >> 4; dal += a_step >> 4; dbh += b_step >> 4; dbl += b_step >> 4;      // start-address field, 16-byte units
            }
            t5_commit(cx.mbar);                            // arrives when every MMA above has finished reading smem / writing TMEM
        }
        mbar_wait(cx.mbar, cx.phase);
        cx.phase ^= 1u;
        TR(112);
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- epilogue: TMEM -> registers (one row of 16 columns per thread), fused bias / relu / mask / dots, store ----
    float v[16];
    float4 mk[4];
    const bool mask_vec = jb.epi == E_MASK && row_ok && j_valid == 16 && (jb.ldmask & 3) == 0 && (jb.sMask & 3) == 0;
    if (mask_vec) {
        const float4* mp = reinterpret_cast<const float4*>(jb.mask + (size_t)z * jb.sMask + (size_t)gi * jb.ldmask + j0);
#pragma unroll
        for (int q = 0; q < 4; ++q) mk[q] = __ldcg(mp + q);
    }
    if (warp < 4) {
        const int nacc = min(kT5Acc, (min(kT5K, K) + 7) >> 3);        // accumulators the first K block initialised
#pragma unroll
        for (int c = 0; c < 16; ++c) v[c] = 0.f;
        for (int a = 0; a < nacc; ++a) {                               // fixed summation order -> deterministic
            unsigned r[16];
            const unsigned taddr = cx.tmem + ((unsigned)(32 * warp) << 16) + (unsigned)(a * kT5N);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                           "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int c = 0; c < 16; ++c) v[c] += __uint_as_float(r[c]);
        }
    }
    if (row_ok) {
        const bool relu = jb.epi == E_BIAS_RELU || jb.epi == E_BIAS_RELU_DOT;
        float dot = 0.f;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const bool ok = j0 + c < N;
            float x = v[c] + s_bias[c];
            if (relu) x = fmaxf(x, 0.f);
            if (jb.epi == E_MASK) {
                float m;
                if (mask_vec) { const float4 q = mk[c >> 2]; m = (c & 3) == 0 ? q.x : (c & 3) == 1 ? q.y : (c & 3) == 2 ? q.z : q.w; }
                else m = ok ? __ldcg(jb.mask + (size_t)z * jb.sMask + (size_t)gi * jb.ldmask + j0 + c) : 0.f;
                x = m > 0.f ? x : 0.f;
            }
            if (!ok) x = 0.f;
            dot = fmaf(x, s_w2[c], dot);
            v[c] = x;
        }
        float* __restrict__ crow = jb.C + (size_t)z * jb.sC + (size_t)gi * jb.ldc + j0;
        if (j_valid == 16 && ((reinterpret_cast<uintptr_t>(crow) & 15) == 0)) {
#pragma unroll
            for (int c = 0; c < 16; c += 4) *reinterpret_cast<float4*>(crow + c) = make_float4(v[c], v[c + 1], v[c + 2], v[c + 3]);
        } else {
#pragma unroll
            for (int c = 0; c < 16; ++c) if (c < j_valid) crow[c] = v[c];
        }
        if (jb.epi == E_BIAS_RELU_DOT) jb.out0[((size_t)z * tn + tj) * jb.i0 + gi] = dot;
        if (J > 0) {     // post-product: P[j] = sum_c v[c] * X[j][c]  (heads off h2, d pi off dh1), see gemm_tile
            for (int j = 0; j < J; ++j) {
                float p = 0.f;
#pragma unroll
                for (int c = 0; c < 16; ++c) p = fmaf(v[c], s_x[j * 17 + c], p);
                jb.out1[(((size_t)z * tn + tj) * J + j) * jb.i0 + gi] = p;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();                           // TMEM and smem are reused by the next tile
    TR(104);
}

// ==========================================================================================================
// bf16 tensor-core path (precision 3): C[128 x 64] = sum_r bf16(A(i,r)) * bf16(B(j,r)), fp32 accumulate in TMEM.
// The BASELINE "Humanoid-shaped, bf16 tensor-core MLP path" config: fp32 master weights / activations / Adam stay as
// they are in L2; every GEMM operand is rounded to bf16 (round-to-nearest-even) on its way into shared memory.
//
//   K is consumed in chunks of 64.  Per chunk: cp.async the fp32 operand tiles into a staging buffer (two in flight),
//   all threads convert staging -> bf16 UMMA operand tiles (always K-major SWIZZLE_NONE, whatever the global layout:
//   the conversion transposes output-contiguous operands), one thread issues 4 tcgen05.mma (M=128, N=64, K=16) and
//   commits to the mbarrier that guards that operand buffer.  Two operand buffers, so the tensor core works on chunk c
//   while chunk c+1 is being converted and chunk c+2 is landing.
//       bf16 K-major canonical layout: byte(o, k) = (o/8)*1024 + (k/8)*128 + (o%8)*16 + (k%8)*2     (LBO 128, SBO 1024)
//   Accumulator: 128 TMEM lanes = 128 rows, 64 fp32 columns; thread t < 128 finishes row t (tcgen05.ld.32x32b).
// ==========================================================================================================
constexpr int kBfM = 128, kBfN = 64, kBfK = 64;
constexpr int kBfPitchRC = kBfK + 4;                                    // staging pitch (floats), reduction-contiguous
constexpr int kBfStageAFloats = kBfM * kBfPitchRC;                      // 8704 (>= 64 * (128 + 4) for the other layout)
constexpr int kBfStageBFloats = kBfN * kBfPitchRC;                      // 4352 (= 64 * (64 + 4))
constexpr int kBfStageBytes = (kBfStageAFloats + kBfStageBFloats) * 4;  // 52,224
constexpr int kBfOpABytes = kBfM * kBfK * 2, kBfOpBBytes = kBfN * kBfK * 2;
constexpr int kBfOpOff = 2 * kBfStageBytes;                             // 104,448
constexpr int kBfEpiOff = kBfOpOff + 2 * (kBfOpABytes + kBfOpBBytes);   // 153,600: bias[64], w2[64]
constexpr int kBfCntOff = kBfEpiOff + 2 * kBfN * 4;                     // 154,112: arrival counters of the two operand buffers
constexpr int kBfMaskOff = kBfCntOff + 16;                              // relu-mask tile [128][68] fp32
static_assert(kBfMaskOff + kBfM * kBfPitchRC * 4 <= kT5PipeBytes, "bf16 tile buffers must fit in the tcgen05 kernel's shared memory");
static_assert(kBfK * (kBfM + 4) <= kBfStageAFloats && kBfK * (kBfN + 4) <= kBfStageBFloats, "staging holds either layout");

__device__ __forceinline__ unsigned bf_idesc() {
    // c_format F32 (1) [4,6) | a_format BF16 (1) [7,10) | b_format BF16 (1) [10,13) | both K-major | N>>3 [17,23) | M>>4 [24,29)
    return (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(kBfN >> 3) << 17) | ((unsigned)(kBfM >> 4) << 24);
}
__device__ __forceinline__ void bf_mma(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned idesc, unsigned acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ unsigned bf_pack(float lo, float hi) {
    unsigned r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));      // first source -> upper half
    return r;
}

// The K loop has no block-wide barrier: every warp OWNS a slice of the tile's output rows (A: 16 rows, B: 8 rows; for
// output-contiguous operands the same slice of columns), stages that slice with its own cp.async groups, converts it and
// then arrives on a shared-memory counter; the warp that arrives last issues the chunk's MMAs and commits them to the
// mbarrier that guards the operand buffer.  Nobody waits for a slower warp except through that mbarrier two chunks later.
//
// fp32 operand slice of warp w (RW = R/8 outputs x 64 reduction elements from (o0, k0)) -> staging.  RC: [R][68]; OC: [64][R+4].
template <bool RC, int R>
__device__ __forceinline__ void bf_stage_w(float* s, const float* __restrict__ g, int ld, int o0, int o_valid, int k0, int k_valid, int vec) {
    constexpr int P = RC ? kBfPitchRC : R + 4;
    constexpr int RW = R / (kThreads / 32);                // outputs owned by a warp: 16 (A) / 8 (B)
    const int lane = threadIdx.x & 31, ow = (threadIdx.x >> 5) * RW;
    if (vec) {
        if (RC) {                        // lane -> (row ow + lane/16 + 2q, 4 k's at (lane%16)*4)
            const int o = ow + (lane >> 4), r = (lane & 15) << 2;
            const int nbk = r < k_valid ? min(4, k_valid - r) * 4 : 0;
            const float* src = g + (size_t)(o0 + o) * ld + k0 + r;
            float* dst = s + o * P + r;
            const size_t step = (size_t)2 * ld;
#pragma unroll
            for (int q = 0; q < RW / 2; ++q) {
                const int nb = (o + 2 * q < o_valid) ? nbk : 0;
                cp_async16(dst + q * 2 * P, nb ? src : g, nb);
                src += step;
            }
        } else {                         // lane -> (k-row lane/VW + (32/VW)q, 4 outputs at ow + (lane%VW)*4)
            constexpr int VW = RW / 4, RS = 32 / VW;       // vectors per k-row of the slice: 4 (A) / 2 (B)
            const int r = lane / VW, o = ow + ((lane % VW) << 2);
            const int nbo = o < o_valid ? min(4, o_valid - o) * 4 : 0;
            const float* src = g + (size_t)(k0 + r) * ld + o0 + o;
            float* dst = s + r * P + o;
            const size_t step = (size_t)RS * ld;
#pragma unroll
            for (int q = 0; q < kBfK / RS; ++q) {
                const int nb = (r + RS * q < k_valid) ? nbo : 0;
                cp_async16(dst + q * RS * P, nb ? src : g, nb);
                src += step;
            }
        }
    } else {            // element-wise path for unaligned operands; the reduction tail is zero-filled to a multiple of 16
        const int k_pad = min(kBfK, (k_valid + 15) & ~15);
#pragma unroll 4
        for (int e = lane; e < RW * kBfK; e += 32) {
            const int o = ow + (RC ? e >> 6 : e % RW), r = RC ? e & 63 : e / RW;
            if (o < o_valid && r < k_pad) {
                float* dst = RC ? s + o * P + r : s + r * P + o;
                if (r < k_valid) cp_async4(dst, RC ? g + (size_t)(o0 + o) * ld + k0 + r : g + (size_t)(k0 + r) * ld + o0 + o);
                else *dst = 0.f;
            }
        }
    }
}

// staging (fp32) -> bf16 K-major UMMA tile, the warp's own slice.  One lane = one 16-byte chunk (8 consecutive k of one
// output row); consecutive lanes take consecutive rows, so a quarter-warp writes one contiguous 128-byte core matrix.
template <bool RC, int R>
__device__ __forceinline__ void bf_convert_w(const float* st, char* op, int k_valid) {
    constexpr int P = RC ? kBfPitchRC : R + 4;
    constexpr int RW = R / (kThreads / 32);
    const int nkc = (min(kBfK, (k_valid + 15) & ~15)) >> 3;
    const int lane = threadIdx.x & 31, ow = (threadIdx.x >> 5) * RW;
    for (int c = lane; c < RW * nkc; c += 32) {
        const int o = ow + c % RW, kc = c / RW;
        float x[8];
        if (RC) {
            const float4 a = *reinterpret_cast<const float4*>(st + o * P + kc * 8);
            const float4 b = *reinterpret_cast<const float4*>(st + o * P + kc * 8 + 4);
            x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
        } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) x[q] = st[(kc * 8 + q) * P + o];
        }
        uint4 pk;
        pk.x = bf_pack(x[0], x[1]); pk.y = bf_pack(x[2], x[3]); pk.z = bf_pack(x[4], x[5]); pk.w = bf_pack(x[6], x[7]);
        *reinterpret_cast<uint4*>(op + (o >> 3) * 1024 + kc * 128 + (o & 7) * 16) = pk;
    }
}

__device__ __forceinline__ void bf_wait(T5Ctx& cx, int b) {
    mbar_wait(b ? cx.mbar_op1 : cx.mbar_op0, (cx.phase_op >> b) & 1u);
    cx.phase_op ^= 1u << b;
}

template <bool A_RC, bool B_RC>
__device__ __forceinline__ void gemm_tile_bf16(const Job& jb, int tile, char* pipe, const RunArgs& ra, T5Ctx& cx) {
    const int M = (jb.rows_dynamic && ra.m_rows > 0) ? min(jb.M, ra.m_rows) : jb.M;
    const int N = jb.N, K = jb.K;
    const int tn = jb.tn;
    const int z = (tile >= jb.tpz) ? 1 : 0;
    const int rem = tile - z * jb.tpz;
    const int ti = rem / tn, tj = rem - ti * tn;
    const int i0 = ti * kBfM, j0 = tj * kBfN;
    if (i0 >= M) return;                                   // uniform per CTA
    const int i_valid = min(kBfM, M - i0), j_valid = min(kBfN, N - j0);
    const float* __restrict__ Ag = jb.A + (size_t)z * jb.sA;
    const float* __restrict__ Bg = jb.B + (size_t)z * jb.sB;
    const int warp = threadIdx.x >> 5;
    float* s_bias = reinterpret_cast<float*>(pipe + kBfEpiOff);
    float* s_w2 = s_bias + kBfN;
    const bool has_bias = jb.epi == E_BIAS || jb.epi == E_BIAS_RELU || jb.epi == E_BIAS_RELU_DOT;

    const int nch = (K + kBfK - 1) / kBfK;
    auto issue = [&](int c) {
        float* sa = reinterpret_cast<float*>(pipe + (c & 1) * kBfStageBytes);
        float* sb = sa + kBfStageAFloats;
        const int k0 = c * kBfK, kv = min(kBfK, K - k0);
        bf_stage_w<A_RC, kBfM>(sa, Ag, jb.lda, i0, i_valid, k0, kv, jb.a_vec);
        bf_stage_w<B_RC, kBfN>(sb, Bg, jb.ldb, j0, j_valid, k0, kv, jb.b_vec);
    };
    issue(0);
    // relu mask of the tile (E_MASK) rides in the first cp.async group: its L2 latency hides behind the K loop
    float* s_mask = reinterpret_cast<float*>(pipe + kBfMaskOff);
    const bool mask_vec = jb.epi == E_MASK && j_valid == kBfN && (jb.ldmask & 3) == 0 && (jb.sMask & 3) == 0;   // tile-uniform
    if (mask_vec) {
        const float* mp = jb.mask + (size_t)z * jb.sMask + (size_t)i0 * jb.ldmask + j0;
#pragma unroll
        for (int q = 0; q < (kBfM * 16) / kThreads; ++q) {
            const int v = threadIdx.x + q * kThreads, o = v >> 4, r = (v & 15) << 2;
            cp_async16(s_mask + o * kBfPitchRC + r, o < i_valid ? mp + (size_t)o * jb.ldmask + r : mp, o < i_valid ? 16 : 0);
        }
    }
    cp_async_commit();
    if (nch > 1) issue(1);
    cp_async_commit();
    // epilogue operands -> shared memory
    if (threadIdx.x < kBfN) s_bias[threadIdx.x] = (has_bias && j0 + threadIdx.x < N) ? __ldcg(jb.bias + (size_t)z * jb.sBias + j0 + threadIdx.x) : 0.f;
    else if (threadIdx.x < 2 * kBfN) {
        const int c = threadIdx.x - kBfN;
        s_w2[c] = (jb.epi == E_BIAS_RELU_DOT && j0 + c < N) ? __ldcg(jb.aux0 + (size_t)z * jb.sAux + j0 + c) : 0.f;
    }
    int* s_cnt = reinterpret_cast<int*>(pipe + kBfCntOff);
    if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
    __syncthreads();                                        // counters zeroed before any warp can arrive

    const unsigned idesc = bf_idesc();
    const int lane = threadIdx.x & 31;
    for (int c = 0; c < nch; ++c) {
        const int b = c & 1;
        const int kv = min(kBfK, K - c * kBfK);
        cp_async_wait<1>();                                 // this lane's copies of chunk c have landed ...
        __syncwarp();                                       // ... and so have the rest of the warp's slice
        if (c >= 2) bf_wait(cx, b);                         // MMAs of chunk c-2 are done with this operand buffer
        const float* sa = reinterpret_cast<const float*>(pipe + b * kBfStageBytes);
        char* oa = pipe + kBfOpOff + b * (kBfOpABytes + kBfOpBBytes);
        char* ob = oa + kBfOpABytes;
        bf_convert_w<A_RC, kBfM>(sa, oa, kv);
        bf_convert_w<B_RC, kBfN>(sa + kBfStageAFloats, ob, kv);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy smem writes -> visible to the tensor core
        __syncwarp();
        if (c + 2 < nch) issue(c + 2);                      // the warp's staging slice just converted is free again
        cp_async_commit();
        if (lane == 0) {
            __threadfence_block();                          // release: the slice is written (and proxy-fenced) before the arrival
            const int seen = atomicAdd(&s_cnt[b], 1);
            if (seen == kThreads / 32 - 1) {                // last warp in: the whole chunk is in place
                s_cnt[b] = 0;                               // next use is chunk c+2, behind the mbarrier wait on this buffer
                __threadfence_block();                      // acquire
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                unsigned long long da = t5_desc(smem_u32(oa), 128u, 1024u, false), db = t5_desc(smem_u32(ob), 128u, 1024u, false);
                const int nk16 = (kv + 15) >> 4;
                for (int k16 = 0; k16 < nk16; ++k16) {
                    bf_mma(cx.tmem, da, db, idesc, (c == 0 && k16 == 0) ? 0u : 1u);
                    da += 16; db += 16;                     // two 128-byte chunk columns per K=16 step (16-byte units)
                }
                t5_commit(b ? cx.mbar_op1 : cx.mbar_op0);
                // the next chunk's MMAs may be issued by another thread (whichever warp arrives last): order them after these
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            }
        }
    }
    TR(205);
    cp_async_wait<0>();
    // consume the outstanding commits in order (the last one covers every MMA of the tile)
    if (nch >= 2) bf_wait(cx, (nch - 2) & 1);
    bf_wait(cx, (nch - 1) & 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    __syncthreads();        // every warp is out of the K loop (the staging buffers become the epilogue's tile buffer)
    TR(206);

    // ---- epilogue, two steps so that all 256 threads work and every global access is coalesced (a thread-per-row
    // epilogue on 4 warps cost ~7,000 cycles per tile):
    //   1. TMEM -> shared memory, raw accumulators.  A warp may only read the TMEM lane quarter (warp % 4), so warps w and
    //      w+4 share rows 32(w%4)..+31 and take 32 columns each (two tcgen05.ld.32x32b.x16, one wait).
    //   2. 16 threads per row (float4 each): + bias, relu / relu-mask, Q-dot partial (shuffle over the 16 lanes), store.
    float* s_out = reinterpret_cast<float*>(pipe);             // [128][68] in the idle staging buffer
    {
        const int lane = threadIdx.x & 31, q4 = warp & 3, half = warp >> 2;
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
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();                           // TMEM and smem are reused by the next tile
    TR(207);
}
