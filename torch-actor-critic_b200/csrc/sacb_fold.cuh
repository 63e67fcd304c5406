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
    // 1. layer-0 weights of this tile's net -> smem [H1][pitchW] (odd pitch: the FFMA loop's column reads hit 32 banks)
    {
        const int n_el = H1 * K0;
        if (pitchW == K0 && ((reinterpret_cast<uintptr_t>(W0g) & 15) == 0)) {
            const int n16 = n_el >> 2;
            for (int v = t; v < n16; v += kThreads) cp_async16(W0s + 4 * v, W0g + 4 * v, 16);
            for (int e = (n16 << 2) + t; e < n_el; e += kThreads) cp_async4(W0s + e, W0g + e);
        } else {
            for (int e = t; e < n_el; e += kThreads) {
                const int n = e / K0, k = e - n * K0;
                cp_async4(W0s + n * pitchW + k, W0g + e);
            }
        }
        cp_async_commit();
    }
    // 2. input rows, transposed: XsT[k][m]
    for (int e = t; e < 32 * ncopy; e += kThreads) {
        const int m = e / ncopy, c = e - m * ncopy;
        XsT[c * 32 + m] = m < i_valid ? __ldcg(Xg + (size_t)(row0 + m) * ldX + c) : 0.f;
    }
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
    const int n = t;                                       // this thread's output column (H1 <= kThreads)
    const float bias = n < H1 ? __ldcg(b0g + n) : 0.f;
    cp_async_wait<0>();
    __syncthreads();
    // 3. h1[m][n] = relu(b0[n] + x[m] . W0[n]) for the tile's 32 rows, fp32 FFMA, k ascending
    const int H1r = (H1 + 7) & ~7;
    if (n < H1r) {
        float acc[32];
#pragma unroll
        for (int m = 0; m < 32; ++m) acc[m] = 0.f;
        if (n < H1) {
            const float* wr = W0s + n * pitchW;
            for (int k = 0; k < K0; ++k) {
                const float w = wr[k];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 x = *reinterpret_cast<const float4*>(XsT + k * 32 + 4 * q);
                    acc[4 * q + 0] = fmaf(x.x, w, acc[4 * q + 0]); acc[4 * q + 1] = fmaf(x.y, w, acc[4 * q + 1]);
                    acc[4 * q + 2] = fmaf(x.z, w, acc[4 * q + 2]); acc[4 * q + 3] = fmaf(x.w, w, acc[4 * q + 3]);
                }
            }
        }
        float* col = stage_a(smem, n >> 6) + (n & 63);
        const bool wr_out = col0 && gwrite && Hout != nullptr && n < H1;
#pragma unroll
        for (int m = 0; m < 32; ++m) {
            const float h = (n < H1 && m < i_valid) ? fmaxf(acc[m] + bias, 0.f) : 0.f;
            col[m * kPitchRC] = h;
            if (wr_out && m < i_valid) Hout[(size_t)(row0 + m) * ldH + n] = h;
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
        const int nst = (H2 + kBK - 1) / kBK;
        for (int s = 0; s < nst; ++s) load_operand<true>(stage_a(smem, s), Ag, jb.lda, i0, i_valid, s * kBK, min(kBK, H2 - s * kBK), jb.a_vec);
        cp_async_commit();
    }
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
    const int n = t;
    const float w2 = n < H2 ? __ldcg(f.c_w2 + (size_t)z * H2 + n) : 0.f;
    cp_async_wait<0>();
    __syncthreads();
    if (col0 && z == 0 && t < 4) {                          // fixed summation order over the tile's rows (deterministic losses)
        float s = 0.f;
        for (int m = 0; m < 32; ++m) s += rowv[m * 4 + t];
        (IS_PI ? f.lparts_pi : f.lparts_q)[(size_t)ti * 4 + t] = s;
    }
    const int H2r = (H2 + 7) & ~7;
    if (n < H2r) {
        float* col = stage_a(smem, n >> 6) + (n & 63);
        const float* dq = dqs + z * 32;
        const bool wr = !IS_PI && col0 && n < H2;
        float cs_b1 = 0.f, cs_w2 = 0.f;
#pragma unroll 8
        for (int m = 0; m < 32; ++m) {
            const float h = col[m * kPitchRC];
            const bool ok = n < H2 && m < i_valid;
            const float d = dq[m];
            const float v = (ok && h > 0.f) ? d * w2 : 0.f;
            col[m * kPitchRC] = v;
            if (wr && m < i_valid) {
                dh2g[(size_t)(i0 + m) * 2 * H2 + (size_t)z * H2 + n] = v;
                cs_b1 += v;
                cs_w2 = fmaf(d, h, cs_w2);
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
        const int nst = (H2 + kBK - 1) / kBK;
        for (int s = 0; s < nst; ++s) load_operand<true>(stage_a(smem, s), jb.A, jb.lda, i0, i_valid, s * kBK, min(kBK, H2 - s * kBK), jb.a_vec);
        cp_async_commit();
    }
    if (t < 256) {
        const int m = t >> 3, a = t & 7, row = i0 + m;
        if (a < A) {
            float gu = 0.f, gls = 0.f;
            if (m < i_valid) {
                const float* eps = ra.eps2 ? ra.eps2 + (size_t)k * ra.eps_stride : nullptr;
                const float glogp = ra.alpha / (float)B;
                const float mu = __ldcg(f.heads + (size_t)row * f.ldh + a), lsr = __ldcg(f.heads + (size_t)row * f.ldh + A + a);
                const float e = eps ? __ldg(eps + (size_t)row * A + a) : 0.f;
                float dpi = 0.f, pp[8];
                for (int t0 = 0; t0 < f.tnH1x2; t0 += 8) {
#pragma unroll
                    for (int u = 0; u < 8; ++u) pp[u] = (t0 + u < f.tnH1x2) ? __ldcg(f.dpart + ((size_t)(t0 + u) * A + a) * B + row) : 0.f;
#pragma unroll
                    for (int u = 0; u < 8; ++u) dpi += pp[u];
                }
                const HeadVals h = head_eval(mu, lsr, e, eps == nullptr, ra);
                const float dcorr = (ra.flags & 1u) ? 2.f * h.th : h.th;
                gu = dpi * ra.act_limit * (1.f - h.th * h.th) + glogp * dcorr;
                const bool inside = (lsr >= ra.log_min_std) && (lsr <= ra.log_max_std);
                gls = inside ? (gu * e * h.std - glogp) : 0.f;
                if (col0) { f.ghead[(size_t)row * f.ldh + a] = gu; f.ghead[(size_t)row * f.ldh + A + a] = gls; }
            }
            ghs[m * 16 + a] = gu; ghs[m * 16 + A + a] = gls;
        }
    }
    const int n = t;
    float wh[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) wh[j] = (j < twoA && n < H2) ? __ldcg(f.a_wh + (size_t)j * H2 + n) : 0.f;
    cp_async_wait<0>();
    __syncthreads();
    if (col0 && t < twoA) {                                 // partial of d loss / d b_heads[t]
        float s = 0.f;
        for (int m = 0; m < 32; ++m) s += ghs[m * 16 + t];
        f.colp_bh[(size_t)ti * 16 + t] = s;
    }
    const int H2r = (H2 + 7) & ~7;
    if (n < H2r) {
        float* col = stage_a(smem, n >> 6) + (n & 63);
        const bool wr = col0 && n < H2;
        float cs = 0.f;
        for (int m = 0; m < 32; ++m) {
            const float h = col[m * kPitchRC];
            float v = 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j)
                if (j < twoA) v = fmaf(ghs[m * 16 + j], wh[j], v);
            v = (n < H2 && m < i_valid && h > 0.f) ? v : 0.f;
            col[m * kPitchRC] = v;
            if (wr && m < i_valid) { adh2g[(size_t)(i0 + m) * H2 + n] = v; cs += v; }
        }
        if (wr) f.colp_ab1[(size_t)ti * H2 + n] = cs;        // partial of d loss / d b1[n] (actor)
    }
    __syncthreads();
}

}  // namespace sacb
