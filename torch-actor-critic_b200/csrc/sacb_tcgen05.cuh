// sacb_tcgen05.cuh -- Blackwell-native GEMM tile: tcgen05.mma (5th-gen tensor cores) with the accumulator in TMEM.
// Included by sacb_device.cuh (inside namespace sacb).
//
//   C[64 x 16] (+epilogue) = sum_r A(i,r) * B(j,r), fp32 storage, 3xTF32 error compensation:
//   every operand tile is split in shared memory into hi = tf32(x) and lo = tf32(x - hi) and the tensor core
//   accumulates a_lo*b_hi + a_hi*b_lo + a_hi*b_hi in fp32 (TMEM) -> fp32-level accuracy.
//
// Operand tiles live in shared memory in the canonical NO-SWIZZLE UMMA layouts (cute/atom/mma_traits_sm100.hpp,
// make_umma_desc), addressed in 16-byte units ("chunks" of 4 tf32):
//   K-major  (reduction contiguous in global memory):  chunk(m, kc) at (m/8)*SBO + kc*LBO + (m%8)*16,  LBO=128, SBO=8192
//   MN-major (output contiguous in global memory):     chunk(k, mc) at (k/8)*LBO + mc*SBO + (k%8)*16,  SBO=128, LBO=(R/4)*128
// so a 16-byte cp.async can drop every global chunk straight into place, and one tcgen05.mma (K = 8 tf32) reads
// 2 K-major chunk columns (start += 256 B per k-step) or one MN-major k-group (start += LBO per k-step).
// Accumulator: M = 64 rows occupy TMEM lanes (m%16) + 32*(m/16) (half of each sub-partition), 16 fp32 columns;
// warp w reads rows 16w..16w+15 with tcgen05.ld.32x32b.x16 (lanes 0..15 hold one row each).
#pragma once

constexpr int kT5M = 64, kT5N = 16, kT5K = 256;                  // tile and K block
constexpr int kT5AHi = 0, kT5ALo = 65536, kT5BHi = 131072, kT5BLo = 147456, kT5PipeBytes = 163840;
constexpr int kT5SboK = 8192, kT5LboK = 128;                     // K-major strides (bytes)
constexpr int kT5SboMN = 128;                                    // MN-major: chunk stride; LBO = (R/4)*128

__device__ __forceinline__ unsigned long long t5_desc(unsigned smem_addr, unsigned lbo_bytes, unsigned sbo_bytes) {
    // cute::UMMA::SmemDescriptor: start[0,14) | LBO[16,30) | SBO[32,46) | version=1 [46,48) | layout_type=SWIZZLE_NONE
    return (unsigned long long)((smem_addr >> 4) & 0x3FFF) | ((unsigned long long)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((unsigned long long)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
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

// chunk loader.  RC = K-major.  R = tile outputs (64 for A, 16 for B).  k_pad = roundup8(k_valid).
template <bool RC>
__device__ __forceinline__ void t5_load(char* buf, const float* __restrict__ g, int ld, int R, int o0, int o_valid,
                                        int k0, int k_valid, int vec) {
    const int k_pad = (k_valid + 7) & ~7;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int l8 = lane & 7, c4 = lane >> 3;
    if (vec) {
        if (RC) {       // unit = 8 rows x 4 chunks; units: (R/8) row groups x ceil(k_pad/16) chunk blocks
            const int MG = R >> 3, CB = (k_pad + 15) >> 4;
            for (int u = warp; u < MG * CB; u += kThreads / 32) {
                const int mg = u % MG, kc = (u / MG) * 4 + c4;
                const int row = mg * 8 + l8, kk = kc * 4;
                if (kk < k_pad) {
                    int nb = 0;
                    if (row < o_valid && kk < k_valid) nb = min(4, k_valid - kk) * 4;
                    const float* src = nb ? g + (size_t)(o0 + row) * ld + k0 + kk : g;
                    cp_async16(buf + mg * kT5SboK + kc * kT5LboK + l8 * 16, src, nb);
                }
            }
        } else {        // unit = 8 k's x 4 chunks; units: (k_pad/8) k groups x (R/16) chunk blocks
            const int KG = k_pad >> 3, CB = R >> 4, lbo = (R >> 2) * 128;
            for (int u = warp; u < KG * CB; u += kThreads / 32) {
                const int kg = u % KG, mc = (u / KG) * 4 + c4;
                const int kk = kg * 8 + l8, oo = mc * 4;
                int nb = 0;
                if (kk < k_valid && oo < o_valid) nb = min(4, o_valid - oo) * 4;
                const float* src = nb ? g + (size_t)(k0 + kk) * ld + o0 + oo : g;
                cp_async16(buf + kg * lbo + mc * kT5SboMN + l8 * 16, src, nb);
            }
        }
    } else {            // element-wise path for unaligned operands; rows beyond o_valid stay stale (feed only unused outputs)
        const int lbo = (R >> 2) * 128;
        const int n_el = o_valid * k_pad;
        for (int e = threadIdx.x; e < n_el; e += kThreads) {
            int o, k;
            if (RC) { o = e / k_pad; k = e - o * k_pad; } else { k = e / o_valid; o = e - k * o_valid; }
            char* dst = RC ? buf + (o >> 3) * kT5SboK + (k >> 2) * kT5LboK + (o & 7) * 16 + (k & 3) * 4
                           : buf + (k >> 3) * lbo + (o >> 2) * kT5SboMN + (k & 7) * 16 + (o & 3) * 4;
            if (k < k_valid) cp_async4(dst, RC ? g + (size_t)(o0 + o) * ld + k0 + k : g + (size_t)(k0 + k) * ld + o0 + o);
            else *reinterpret_cast<float*>(dst) = 0.f;
        }
    }
}

// hi/lo split of everything the loader touched: x -> hi (in place), lo -> buf_lo (same layout).
template <bool RC>
__device__ __forceinline__ void t5_split(char* hi, char* lo, int R, int k_valid) {
    const int k_pad = (k_valid + 7) & ~7;
    int groups, per_group, gstride;                   // float4 units
    if (RC) { groups = R >> 3; per_group = (k_pad >> 2) * 8; gstride = kT5SboK / 16; }
    else { groups = k_pad >> 3; per_group = (R >> 2) * 8; gstride = (R >> 2) * 8; }
    const int total = groups * per_group;
    for (int i = threadIdx.x; i < total; i += kThreads) {
        const int gi = i / per_group, off = (gi * gstride + (i - gi * per_group)) * 16;
        float4 x = *reinterpret_cast<const float4*>(hi + off);
        float4 h, l;
        h.x = __uint_as_float(f2tf32(x.x)); l.x = __uint_as_float(f2tf32(x.x - h.x));
        h.y = __uint_as_float(f2tf32(x.y)); l.y = __uint_as_float(f2tf32(x.y - h.y));
        h.z = __uint_as_float(f2tf32(x.z)); l.z = __uint_as_float(f2tf32(x.z - h.z));
        h.w = __uint_as_float(f2tf32(x.w)); l.w = __uint_as_float(f2tf32(x.w - h.w));
        *reinterpret_cast<float4*>(hi + off) = h;
        *reinterpret_cast<float4*>(lo + off) = l;
    }
}

struct T5Ctx { unsigned tmem; unsigned mbar; unsigned phase; };   // TMEM base address, mbarrier smem address, its parity

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

    const unsigned idesc = t5_idesc(!A_RC, !B_RC);
    const unsigned a_lbo = A_RC ? kT5LboK : (kT5M / 4) * 128, a_sbo = A_RC ? kT5SboK : kT5SboMN;
    const unsigned b_lbo = B_RC ? kT5LboK : (kT5N / 4) * 128, b_sbo = B_RC ? kT5SboK : kT5SboMN;
    const unsigned a_step = A_RC ? 256 : a_lbo, b_step = B_RC ? 256 : b_lbo;      // bytes per k-step of 8

    // epilogue operands for this thread's row (warps 0..3, lanes 0..15: row = 16*warp + lane)
    const int erow = 16 * warp + lane;
    const bool e_thread = warp < 4 && lane < 16;
    const int gi = i0 + erow;
    const bool row_ok = e_thread && gi < M;

    for (int kb = 0; kb < K; kb += kT5K) {
        const int kv = min(kT5K, K - kb);
        t5_load<A_RC>(a_hi, Ag, jb.lda, kT5M, i0, i_valid, kb, kv, jb.a_vec);
        t5_load<B_RC>(b_hi, Bg, jb.ldb, kT5N, j0, j_valid, kb, kv, jb.b_vec);
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        TR(110);
        t5_split<A_RC>(a_hi, a_lo, kT5M, kv);
        t5_split<B_RC>(b_hi, b_lo, kT5N, kv);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy smem writes -> visible to the tensor core
        __syncthreads();
        TR(111);
        if (threadIdx.x == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const unsigned ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
            const int nk8 = (kv + 7) >> 3;
            for (int k8 = 0; k8 < nk8; ++k8) {
                const unsigned long long dah = t5_desc(ah + k8 * a_step, a_lbo, a_sbo), dal = t5_desc(al + k8 * a_step, a_lbo, a_sbo);
                const unsigned long long dbh = t5_desc(bh + k8 * b_step, b_lbo, b_sbo), dbl = t5_desc(bl + k8 * b_step, b_lbo, b_sbo);
                t5_mma(cx.tmem, dal, dbh, idesc, (kb | k8) ? 1u : 0u);
                t5_mma(cx.tmem, dah, dbl, idesc, 1u);
                t5_mma(cx.tmem, dah, dbh, idesc, 1u);
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
    if (warp < 4) {
        unsigned r[16];
        const unsigned taddr = cx.tmem + ((unsigned)(32 * warp) << 16);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                       "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int c = 0; c < 16; ++c) v[c] = __uint_as_float(r[c]);
    }
    if (row_ok) {
        const bool has_bias = jb.epi == E_BIAS || jb.epi == E_BIAS_RELU || jb.epi == E_BIAS_RELU_DOT;
        const bool relu = jb.epi == E_BIAS_RELU || jb.epi == E_BIAS_RELU_DOT;
        float dot = 0.f;
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const int gj = j0 + c;
            const bool ok = gj < N;
            float x = v[c];
            if (has_bias && ok) x += __ldcg(jb.bias + (size_t)z * jb.sBias + gj);
            if (relu) x = fmaxf(x, 0.f);
            if (jb.epi == E_MASK) x = (ok && __ldcg(jb.mask + (size_t)z * jb.sMask + (size_t)gi * jb.ldmask + gj) > 0.f) ? x : 0.f;
            if (!ok) x = 0.f;
            if (jb.epi == E_BIAS_RELU_DOT && ok) dot = fmaf(x, __ldcg(jb.aux0 + (size_t)z * jb.sAux + gj), dot);
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
        const int J = jb.i1;
        if (J > 0) {     // post-product: P[j] = sum_c v[c] * X[j][c]  (heads off h2, d pi off dh1), see gemm_tile
            const float* __restrict__ X = jb.aux1 + (size_t)z * jb.sAux;
            for (int j = 0; j < J; ++j) {
                float p = 0.f;
#pragma unroll
                for (int c = 0; c < 16; ++c)
                    if (j0 + c < N) p = fmaf(v[c], __ldcg(jb.i3 ? X + (size_t)(j0 + c) * jb.i2 + j : X + (size_t)j * jb.i2 + j0 + c), p);
                jb.out1[(((size_t)z * tn + tj) * J + j) * jb.i0 + gi] = p;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();                           // TMEM and smem are reused by the next tile
    TR(104);
}
