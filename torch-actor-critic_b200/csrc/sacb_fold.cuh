// sacb_fold.cuh -- A-operand generators of the FOLDED update program (included by sacb_device.cuh).
//
// The first persistent program ran one SAC update (reference sac/algorithm.py:275-278) as 16 grid-synchronised phases; six
// of them were thin -- the K = 17 / 23 input-layer GEMMs and the per-row jobs (tanh-Gaussian heads, Bellman backup,
// min-Q routing, head gradients) -- and cost ~3.7 us each (an L2 round trip + a grid barrier) for microseconds of math.
// Here the consumer GEMM tile computes those values itself, straight into the shared-memory stage buffers of its A operand
// (K <= 256, so the whole reduction is resident), redundantly across the 8-16 column tiles of a row block:
//
//   GEN_L0     A[m][n] = relu(b0[n] + sum_k X[m][k] W0[n][k])      layer 0 of actor / critic / target critic in fp32 FFMA
//              (networks/linear.py:33-35,62-65); for the target critic and the critic on (s, pi) the action columns of X are
//              the tanh-Gaussian sample itself (networks/linear.py:37-51), evaluated from the head partial sums
//   GEN_BACKUP A[m][n] = dq_z[m] W2_z[n] [h2_z[m][n] > 0]           backup, twin MSE, dL/dq (sac/algorithm.py:63-72)
//   GEN_PI     same with dq from the min-Q routing of the actor loss (sac/algorithm.py:40-41)
//   GEN_GHEAD  A[m][n] = (sum_j g[m][j] Wh[j][n]) [ah2[m][n] > 0]   head gradients and the back-projection through the heads
//
// The tiles of column 0 also store what later phases need from global memory (layer-0 activations, dh2, log-probs, the
// per-row diagnostics the parity tests read) and the per-row-block column sums that become the bias / W2 gradients.
// One update = 8 grid-synchronised phases (10 with a gradient exchange): see build_fold_program in sacb_program.inl.
#pragma once

namespace sacb {

__device__ __forceinline__ float* stage_a(float* smem, int s) { return smem + s * kStageFloats; }

// [mu | log_std_raw] of (row, a) from the per-column-tile partials of the A.L1 post-product; same summation order as
// row_heads_tile (groups of 8, bias first)
__device__ __forceinline__ void heads_from_partials(const FoldPtrs& f, int A, int B, int row, int a, float& mu, float& lsr) {
    mu = __ldcg(f.a_bh + a); lsr = __ldcg(f.a_bh + A + a);
    float pm[8], pl[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const bool ok = u < f.tnH2;
        pm[u] = ok ? __ldcg(f.hpart + ((size_t)u * 2 * A + a) * B + row) : 0.f;
        pl[u] = ok ? __ldcg(f.hpart + ((size_t)u * 2 * A + A + a) * B + row) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) { mu += pm[u]; lsr += pl[u]; }
}

// ----------------------------------------------------------------------------------------------------------
// GEN_L0
// ----------------------------------------------------------------------------------------------------------
// shared-memory address of the generators' mbarrier (the tcgen05 path's first mbarrier slot; unused by this kernel variant)
__device__ __forceinline__ unsigned gen_mbar_addr(float* smem) {
    return smem_u32(reinterpret_cast<char*>(smem) + (kPipeFloats + kGenFloats) * 4 + kJobsBytes + kMaxPhases * (int)sizeof(Phase));
}

// Register tile of the generators' small products: thread t owns rows 8*(t/64) .. +7 and columns (t%64) + 64*j, j = 0..3
// (j = the stage buffer the column lives in), so the row operands are warp-wide broadcasts and the column operands are
// conflict-free, and every shared-memory value is reused 4x / 8x from registers.

__device__ __forceinline__ void gen_l0(const Job& jb, const RunArgs& ra, float* smem, int i0, int i_valid, int z, bool col0, int kk) {
    const FoldPtrs& f = ra.f;
    float* gsm = smem + kPipeFloats;
    float* XsT = gsm + kGenXT;
    float* W0s = gsm + kGenW0;
    const int t = threadIdx.x;
    // (job fields live in shared memory: copy what the loops use into registers, stores through float* could alias them)
    const int src = jb.gsrc, O = ra.O, A = ra.A, B = ra.B, gwrite = jb.gwrite;
    const int K0 = src == 0 ? O : O + A;
    const int H1 = jb.K;
    const int pitchW = K0 | 1;
    const size_t par = (size_t)(kk & 1);
    const int row0 = jb.grow0 + i0;
    const float* __restrict__ W0g; const float* __restrict__ b0g; const float* __restrict__ Xg; int ldX, ncopy;
    float* Hout = nullptr; int ldH = 0;
    switch (src) {
        case 0: W0g = f.a_w0; b0g = f.a_b0; Xg = f.s2 + par * f.s2_stride; ldX = f.lds; ncopy = O; Hout = f.ah1; ldH = H1; break;
        case 1: W0g = f.t_w0 + (size_t)z * H1 * K0; b0g = f.t_b0 + z * H1; Xg = f.s2 + par * f.s2_stride; ldX = f.lds; ncopy = O; break;
        case 2: W0g = f.c_w0 + (size_t)z * H1 * K0; b0g = f.c_b0 + z * H1; Xg = f.xs + par * f.xs_stride; ldX = f.ldx; ncopy = O;
                Hout = f.cp1 + z * H1; ldH = 2 * H1; break;
        default: W0g = f.c_w0 + (size_t)z * H1 * K0; b0g = f.c_b0 + z * H1; Xg = f.xs + par * f.xs_stride; ldX = f.ldx; ncopy = K0;
                Hout = f.ch1 + z * H1; ldH = 2 * H1; break;
    }
    // 1. layer-0 weights of this tile's net -> smem [H1][pitchW] (odd pitch: the FFMA loop's column reads hit 32 banks).
    //    The matrix is one contiguous block of the arena: ONE TMA bulk copy (cp.async.bulk, SASS UBLKCP) issued by one thread,
    //    completion on an mbarrier -- instead of 6 LDGSTS per thread.  Unaligned / even-K0 shapes fall back to cp.async.
    const int n_el = H1 * K0;
    const bool bulk = pitchW == K0 && ((reinterpret_cast<uintptr_t>(W0g) & 15) == 0) && ((n_el & 3) == 0);
    const unsigned mbar = gen_mbar_addr(smem);
    if (bulk) {
        if (t == 0) {
            const unsigned bytes = (unsigned)n_el * 4u;
            mbar_init(mbar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // earlier generic-proxy reads of W0s are done (block barrier)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(W0s)), "l"(W0g), "r"(bytes), "r"(mbar) : "memory");
        }
    } else {
        for (int e = t; e < n_el; e += kThreads) {
            const int n = e / K0, k = e - n * K0;
            cp_async4(W0s + n * pitchW + k, W0g + e);
        }
        cp_async_commit();
    }
    TR(130);
    // 2. input rows, transposed: XsT[k][m] (all loads in flight before the first store)
    {
        float xv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = t + u * kThreads;
            const int m = e / ncopy, c = e - m * ncopy;
            xv[u] = (e < 32 * ncopy && m < i_valid) ? __ldcg(Xg + (size_t)(row0 + m) * ldX + c) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = t + u * kThreads;
            const int m = e / ncopy, c = e - m * ncopy;
            if (e < 32 * ncopy) XsT[c * 32 + m] = xv[u];
        }
    }
    TR(131);
    // 2b. action columns = the policy's own sample (networks/linear.py:37-51): pass 0 (a2, eps1) for the target critic,
    //     pass 1 (pi, eps2) for the critic of the actor loss.  8 lanes per row, lane a < A.
    if (src == 1 || src == 2) {
        const int pass = src == 2 ? 1 : 0;
        const float* eps = pass ? ra.eps2 : ra.eps1;
        if (eps) eps += (size_t)kk * ra.eps_stride;
        if (t < 256) {
            const int m = t >> 3, a = t & 7, row = row0 + m;
            const bool active = m < i_valid && a < A;
            float sl = 0.f, sc = 0.f, act = 0.f, mu = 0.f, lsr = 0.f;
            if (active) {
                heads_from_partials(f, A, B, row, a, mu, lsr);
                const float e = eps ? __ldg(eps + (size_t)row * A + a) : 0.f;
                const HeadVals h = head_eval(mu, lsr, e, eps == nullptr, ra);
                act = __fmul_rn(h.th, ra.act_limit);
                sl = h.logp_term; sc = h.corr_term;
            }
            if (a < A) XsT[(O + a) * 32 + m] = act;
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {
                sl += __shfl_xor_sync(0xffffffffu, sl, o);
                sc += __shfl_xor_sync(0xffffffffu, sc, o);
            }
            if (col0 && z == 0 && m < i_valid) {
                if (a == 0) f.logp[pass * B + row] = __fsub_rn(sl, sc);
                if (a < A) {
                    f.heads[(size_t)row * f.ldh + a] = mu;            // [mu | log_std_raw] for the head-gradient rows
                    f.heads[(size_t)row * f.ldh + A + a] = lsr;
                    (pass ? f.dbg_pi : f.dbg_a2)[(size_t)row * f.ldx + a] = act;
                }
            }
        }
    }
    TR(132);
    const int cg = t & 63, rg = t >> 6;                     // this thread's column group / row group (kThreads == 256)
    float bias[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) bias[j] = (cg + 64 * j < H1) ? __ldcg(b0g + cg + 64 * j) : 0.f;
    if (!bulk) cp_async_wait<0>();
    __syncthreads();                                        // X^T complete; the mbarrier's init is visible to every thread
    if (bulk) mbar_wait(mbar, 0);
    TR(133);
    // 3. h1[m][n] = relu(b0[n] + x[m] . W0[n]) for the tile's 32 rows, fp32 FFMA, k ascending
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    {
        // columns beyond H1 read (finite) stale shared memory and are masked below
        const float* w0 = W0s + cg * pitchW;
        const int cstride = 64 * pitchW;
        const float* xr = XsT + 8 * rg;
#pragma unroll 1
        for (int k = 0; k < K0; ++k) {
            const float4 xa = *reinterpret_cast<const float4*>(xr + k * 32);
            const float4 xb = *reinterpret_cast<const float4*>(xr + k * 32 + 4);
            const float x[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
            float w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) w[j] = (cg + 64 * j < H1) ? w0[j * cstride + k] : 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(x[i], w[j], acc[i][j]);
        }
    }
    TR(134);
    const bool wr_out = col0 && gwrite && Hout != nullptr;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int n = cg + 64 * j;
        float* col = stage_a(smem, j) + cg;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int m = 8 * rg + i;
            const float h = (n < H1 && m < i_valid) ? fmaxf(acc[i][j] + bias[j], 0.f) : 0.f;
            col[m * kPitchRC] = h;
            if (wr_out && n < H1 && m < i_valid) Hout[(size_t)(row0 + m) * ldH + n] = h;
        }
    }
    __syncthreads();
}

// ----------------------------------------------------------------------------------------------------------
// GEN_BACKUP / GEN_PI: the loss rows (row_loss_tile) computed per tile, dh2 generated in place over the relu-mask tile
// ----------------------------------------------------------------------------------------------------------
template <bool IS_PI>
__device__ __forceinline__ void gen_loss(const Job& jb, const RunArgs& ra, float* smem, int i0, int i_valid, int z, int ti, bool col0, int k) {
    const FoldPtrs& f = ra.f;
    float* misc = smem + kPipeFloats + kGenMisc;
    float* dqs = misc;                                      // [2][32]
    const int t = threadIdx.x;
    const int H2 = jb.K, B = ra.B, nt = f.tnH2;
    float* const dh2g = f.dh2; float* const colp_b1 = f.colp_b1; float* const colp_w2 = f.colp_w2;
    const size_t par = (size_t)(k & 1);
    // relu-mask tile (forward activations of layer 1, this z) -> the A stage buffers
    {
        const float* Ag = jb.A + (size_t)z * jb.sA;
        const int lda = jb.lda, a_vec = jb.a_vec;
        const int nst = (H2 + kBK - 1) / kBK;
        for (int s = 0; s < nst; ++s) load_operand<true>(stage_a(smem, s), Ag, lda, i0, i_valid, s * kBK, min(kBK, H2 - s * kBK), a_vec);
        cp_async_commit();
    }
    TR(120);
    // per-row scalars: 8 lanes per row (lane tt sums partial tile tt of every net), all loads of the tile in flight at once
    const float Bf = (float)B;
    float* rowv = misc + 64;                                // [32][4] per-row loss terms
    if (t < 256) {
        const int m = t >> 3, tt = t & 7, row = i0 + m;
        const bool rok = m < i_valid;
        float p1 = 0.f, p2 = 0.f, p3 = 0.f, p4 = 0.f;
        if (rok && tt < nt) {
            const float* qp = IS_PI ? f.qppart : f.qpart;
            p1 = __ldcg(qp + (size_t)tt * B + row);
            p2 = __ldcg(qp + ((size_t)nt + tt) * B + row);
            if (!IS_PI) {
                p3 = __ldcg(f.qtpart + (size_t)tt * B + row);
                p4 = __ldcg(f.qtpart + ((size_t)nt + tt) * B + row);
            }
        }
        const float b2a = __ldcg(f.c_b2), b2b = __ldcg(f.c_b2 + 1);
        float bt0 = 0.f, bt1 = 0.f, rw = 0.f, dn = 0.f, lp = 0.f;
        if (rok) {
            if (!IS_PI) {
                bt0 = __ldcg(f.t_b2); bt1 = __ldcg(f.t_b2 + 1);
                rw = __ldcg(f.rd + par * f.rd_stride + row); dn = __ldcg(f.rd + par * f.rd_stride + B + row);
                lp = __ldcg(f.logp + row);
            } else {
                lp = __ldcg(f.logp + B + row);
            }
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            p1 += __shfl_xor_sync(0xffffffffu, p1, o); p2 += __shfl_xor_sync(0xffffffffu, p2, o);
            p3 += __shfl_xor_sync(0xffffffffu, p3, o); p4 += __shfl_xor_sync(0xffffffffu, p4, o);
        }
        if (tt == 0) {
            float dq1 = 0.f, dq2 = 0.f, v0 = 0.f, v1 = 0.f;
            if (rok) {
                const float q1 = p1 + b2a, q2 = p2 + b2b;
                if (!IS_PI) {
                    const float q1t = p3 + bt0, q2t = p4 + bt1;
                    const float qt = fminf(q1t, q2t);
                    // backup = reward_scale*r + gamma*(1-d)*(qt - alpha*logp)          sac/algorithm.py:65-67
                    const float backup = __fadd_rn(__fmul_rn(ra.reward_scale, rw),
                                                   __fmul_rn(__fmul_rn(ra.gamma, __fsub_rn(1.f, dn)), __fsub_rn(qt, __fmul_rn(ra.alpha, lp))));
                    const float e1 = __fsub_rn(q1, backup), e2 = __fsub_rn(q2, backup);
                    dq1 = 2.f * e1 / Bf; dq2 = 2.f * e2 / Bf;
                    v0 = e1 * e1; v1 = e2 * e2;
                    if (col0 && z == 0) {
                        float* qb = f.qbuf;
                        qb[row] = q1; qb[B + row] = q2; qb[2 * B + row] = q1t; qb[3 * B + row] = q2t; qb[4 * B + row] = backup;
                        qb[5 * B + row] = dq1; qb[6 * B + row] = dq2;
                    }
                } else {
                    const float qmin = fminf(q1, q2), gq = -1.f / Bf;
                    dq1 = q1 < q2 ? gq : (q1 == q2 ? 0.5f * gq : 0.f);      // torch.min backward: ties split
                    dq2 = q2 < q1 ? gq : (q1 == q2 ? 0.5f * gq : 0.f);
                    v0 = __fsub_rn(__fmul_rn(ra.alpha, lp), qmin);         // alpha*logp - q_pi      sac/algorithm.py:41
                    if (col0 && z == 0) {
                        float* qb = f.qbuf;
                        qb[7 * B + row] = q1; qb[8 * B + row] = q2; qb[5 * B + row] = dq1; qb[6 * B + row] = dq2;
                    }
                }
            }
            dqs[m] = dq1; dqs[32 + m] = dq2;
            rowv[m * 4 + 0] = v0; rowv[m * 4 + 1] = v1; rowv[m * 4 + 2] = IS_PI ? 0.f : dq1; rowv[m * 4 + 3] = IS_PI ? 0.f : dq2;
        }
    }
    TR(121);
    const int n = t;
    const float w2 = n < H2 ? __ldcg(f.c_w2 + (size_t)z * H2 + n) : 0.f;
    cp_async_wait<0>();
    __syncthreads();
    TR(122);
    if (col0 && z == 0 && t < 4) {                          // fixed summation order over the tile's rows (deterministic losses)
        float s = 0.f;
#pragma unroll 4
        for (int m = 0; m < 32; ++m) s += rowv[m * 4 + t];
        (IS_PI ? f.lparts_pi : f.lparts_q)[(size_t)ti * 4 + t] = s;
    }
    const int H2r = (H2 + 7) & ~7;
    if (n < H2r) {
        float* col = stage_a(smem, n >> 6) + (n & 63);
        const bool wr = !IS_PI && col0 && n < H2;
        // Rolled on purpose: a fully unrolled version of this loop (650 instructions) measured 6.3k cycles against 3.4k.
        // 4 rows per iteration, every shared-memory load before the first store (col[] and dq[] may alias for the compiler).
        float cs_b1 = 0.f, cs_w2 = 0.f;
        const float* dqz = dqs + z * 32;
#pragma unroll 1
        for (int m0 = 0; m0 < 32; m0 += 4) {
            const float4 d4 = *reinterpret_cast<const float4*>(dqz + m0);
            const float d[4] = {d4.x, d4.y, d4.z, d4.w};
            float h[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) h[u] = col[(m0 + u) * kPitchRC];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int m = m0 + u;
                const bool ok = n < H2 && m < i_valid;
                const float v = (ok && h[u] > 0.f) ? d[u] * w2 : 0.f;
                col[m * kPitchRC] = v;
                if (wr && m < i_valid) {
                    dh2g[(size_t)(i0 + m) * 2 * H2 + (size_t)z * H2 + n] = v;
                    cs_b1 += v;
                    cs_w2 = fmaf(d[u], h[u], cs_w2);
                }
            }
        }
        if (wr) {
            colp_b1[(size_t)ti * 2 * H2 + (size_t)z * H2 + n] = cs_b1;        // partial of d loss / d b1[z][n]
            colp_w2[(size_t)ti * 2 * H2 + (size_t)z * H2 + n] = cs_w2;        // partial of d loss / d W2[z][n]
        }
    }
    __syncthreads();
}

// ----------------------------------------------------------------------------------------------------------
// GEN_GHEAD: head-gradient rows (row_ghead_tile) + the back-projection dh2 = (g . Wh) * relu'(h2)
// ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void gen_ghead(const Job& jb, const RunArgs& ra, float* smem, int i0, int i_valid, int ti, bool col0, int k) {
    const FoldPtrs& f = ra.f;
    float* ghs = smem + kPipeFloats + kGenMisc;            // [32][16]
    const int t = threadIdx.x;
    const int H2 = jb.K, A = ra.A, B = ra.B, twoA = 2 * A;
    float* const adh2g = f.adh2;
    {
        const float* Ag = jb.A;
        const int lda = jb.lda, a_vec = jb.a_vec;
        const int nst = (H2 + kBK - 1) / kBK;
        for (int s = 0; s < nst; ++s) load_operand<true>(stage_a(smem, s), Ag, lda, i0, i_valid, s * kBK, min(kBK, H2 - s * kBK), a_vec);
        cp_async_commit();
    }
    const int cg = t & 63, rg = t >> 6;
    // head weights of this thread's 4 columns: wh[j][c] = Wh[j][cg + 64 c]
    float wh[16][4];
#pragma unroll
    for (int j = 0; j < 16; ++j)
#pragma unroll
        for (int c = 0; c < 4; ++c) wh[j][c] = (j < twoA && cg + 64 * c < H2) ? __ldcg(f.a_wh + (size_t)j * H2 + cg + 64 * c) : 0.f;
    if (t < 256) {
        const int m = t >> 3, a = t & 7, row = i0 + m;
        if (a < A) {
            float gu = 0.f, gls = 0.f;
            if (m < i_valid) {
                const float* eps = ra.eps2 ? ra.eps2 + (size_t)k * ra.eps_stride : nullptr;
                const float glogp = ra.alpha / (float)B;
                const float mu = __ldcg(f.heads + (size_t)row * f.ldh + a), lsr = __ldcg(f.heads + (size_t)row * f.ldh + A + a);
                const float e = eps ? __ldg(eps + (size_t)row * A + a) : 0.f;
                float pp[16];                               // d pi partials: every load in flight at once, summed in tile order
#pragma unroll
                for (int u = 0; u < 16; ++u) pp[u] = (u < f.tnH1x2) ? __ldcg(f.dpart + ((size_t)u * A + a) * B + row) : 0.f;
                float dpi = 0.f;
#pragma unroll
                for (int u = 0; u < 16; ++u) dpi += pp[u];
                const HeadVals h = head_eval(mu, lsr, e, eps == nullptr, ra);
                const float dcorr = (ra.flags & 1u) ? 2.f * h.th : h.th;
                gu = dpi * ra.act_limit * (1.f - h.th * h.th) + glogp * dcorr;
                const bool inside = (lsr >= ra.log_min_std) && (lsr <= ra.log_max_std);
                gls = inside ? (gu * e * h.std - glogp) : 0.f;
                if (col0) { f.ghead[(size_t)row * f.ldh + a] = gu; f.ghead[(size_t)row * f.ldh + A + a] = gls; }
            }
            ghs[m * 16 + a] = gu; ghs[m * 16 + A + a] = gls;
        }
        if (a == 7)
            for (int j = twoA; j < 16; ++j) ghs[m * 16 + j] = 0.f;    // unused tail of the row: finite (it meets zero weights)
    }
    TR(141);
    cp_async_wait<0>();
    __syncthreads();
    TR(142);
    if (col0 && t < twoA) {                                 // partial of d loss / d b_heads[t]
        float s = 0.f;
        for (int m = 0; m < 32; ++m) s += ghs[m * 16 + t];
        f.colp_bh[(size_t)ti * 16 + t] = s;
    }
    // dh2 for rows 8*rg .. +7, columns cg + 64c: g rows are warp-wide broadcasts reused for 4 columns from registers.
    // One row per iteration, rolled (a fully unrolled version measured slower).
    float cs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
    for (int i = 0; i < 8; ++i) {
        const int m = 8 * rg + i;
        const float* gr = ghs + m * 16;
        const float4 g0 = *reinterpret_cast<const float4*>(gr), g1 = *reinterpret_cast<const float4*>(gr + 4),
                     g2 = *reinterpret_cast<const float4*>(gr + 8), g3 = *reinterpret_cast<const float4*>(gr + 12);
        const float g[16] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w, g2.x, g2.y, g2.z, g2.w, g3.x, g3.y, g3.z, g3.w};
        float hv[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) hv[c] = (stage_a(smem, c) + cg)[m * kPitchRC];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float a = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) a = fmaf(g[j], wh[j][c], a);   // j ascending; entries beyond 2A are zero
            const int n = cg + 64 * c;
            const float o = (n < H2 && m < i_valid && hv[c] > 0.f) ? a : 0.f;
            (stage_a(smem, c) + cg)[m * kPitchRC] = o;
            if (col0 && n < H2 && m < i_valid) { adh2g[(size_t)(i0 + m) * H2 + n] = o; cs[c] += o; }
        }
    }
    if (col0) {
        // partial of d loss / d b1[n] (actor): the 4 row groups of a column meet in shared memory, fixed order
        float* csum = smem + kPipeFloats + kGenXT;          // [4][256] (the X^T scratch is idle in this generator)
#pragma unroll
        for (int c = 0; c < 4; ++c) csum[rg * 256 + cg + 64 * c] = cs[c];
    }
    __syncthreads();
    if (col0 && t < H2) {
        const float* csum = smem + kPipeFloats + kGenXT;
        f.colp_ab1[(size_t)ti * H2 + t] = (csum[t] + csum[256 + t]) + (csum[512 + t] + csum[768 + t]);
    }
}

}  // namespace sacb
