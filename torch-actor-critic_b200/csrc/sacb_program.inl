// sacb_program.inl -- the phase programs (included by sacb_engine.cu after Builder / sacb_engine are defined).
//
// Program of ONE SAC gradient update (reference sac/algorithm.py:275-278).  With a gradient exchange between backward
// and optimizer step (world > 1) it has 18 grid-synchronised phases:
//
//   P0*  gather(k)                                   (* first update of a launch only; later batches are prefetched)
//   ---- update_critic (sac/algorithm.py:115-141) ----
//   P1   A.L0 (s' -> ah1)                  ||  C.L0[q1] ([s|a] -> ch1)
//   P2   A.L1 (+ heads as post-product)    ||  C.L0[q2]
//   P3   tanh-Gaussian rows (a2, logp | pi, logp)  ||  C.L1[q1] (+ Q-dot)  ||  C.L1[q2], first row blocks
//   P4   T.L0 twin ([s'|a2] -> th1)
//   P5   T.L1 twin (+ Q-dot)               ||  C.L1[q2], remaining row blocks
//   P6   backup / dq / dh2 rows
//   P7   C.dh1 twin  ||  db1  ||  dW2  ||  loss_q, db2
//   P8   C.dW1 twin  ||  C.dW0  ||  db0
//   P9   [exchange]  Adam(critic) + Polyak
//   ---- update_policy (sac/algorithm.py:143-162) ----
//   P10  Cpi.L0 twin ([s|pi] -> ch1)
//   P11  Cpi.L1 twin (+ Q-dot)
//   P12  min-Q rows (loss_pi parts, dq routing, dh2)
//   P13  Cpi.dh1 twin (+ d pi as post-product over the action columns of W0)
//   P14  head-gradient rows (g_mu, g_ls)
//   P15  A.dh2 || loss_pi
//   P16  A.dh1
//   P17  A.dWh || dbh || A.dW1 || db1 || A.dW0 || db0        (every K = batch reduction in one phase)
//   P18  [exchange]  Adam(actor)  ||  gather(k+1)
//
// On ONE GPU (mma.sync / FFMA tiles) the Adam step and the Polyak move are applied by the jobs that finish a gradient, in
// their epilogues (Job.adam), P9 and P18 disappear (16 phases) and the actor's backward becomes
//   P15' A.dh2 || loss_pi      P16' A.dWh+Adam || dbh+Adam || A.dh1      P17' A.dW1+Adam || db1 || A.dW0+Adam || db0 || gather(k+1)
// (a weight matrix is stepped one phase after the dX GEMM that reads it; the s' staging tile is parity double-buffered
// because the gather of the next batch now overlaps A.dW0, which still reads this batch's s').
//
// The two post-products need 2A <= 16 resp. A <= 16; wider action spaces, and the bf16 path, use a separate heads GEMM /
// d-pi GEMM phase (generic path, exercised by the Humanoid-shaped tests).

static void build_programs(sacb_engine* e) {
    e->jobs.clear();
    e->phases.clear();
    Builder b(e);
    const sacb_config& c = e->cfg;
    const sacb_layout_info& L = e->lay;
    const int O = c.obs_dim, A = c.act_dim, H1 = c.hidden1, H2 = c.hidden2, B = c.batch_size;
    const int OA = O + A, ldx = e->ldx, lds = e->lds, ldh = e->ldh, lddpi = e->lddpi;
    float *PA = e->pa, *PC = e->pc, *PT = e->pt, *GA = e->ga, *GC = e->gc;
    const long long sW1 = L.critic_w1[1] - L.critic_w1[0];
    const int row_tiles = e->row_tiles;
    // bf16 path: every matrix product is a GEMM job of its own (bf16 operands), nothing rides in an fp32 epilogue
    const bool fuse_heads = 2 * A <= 16 && e->tf32x3 != 3, fuse_dpi = A <= 16 && e->tf32x3 != 3;
    const int tnH1 = (H1 + e->tile_n - 1) / e->tile_n;
    const bool actor_sync = !(c.flags & SACB_FLAG_NO_ACTOR_GRAD_SYNC);
    // SACB_FLAG_PI_FROM_STATES: the canonical actor loss, pi = actor(states) instead of actor(next_states)
    // (sac/algorithm.py:37).  The trunk can then no longer be shared between the two noise draws: update_policy evaluates
    // the actor again on the state columns of xs = [s | a], and the next batch's gather (which overwrites xs) can no longer
    // overlap the last phase, so it runs as the first phase of every update.
    const bool pi_states = (c.flags & SACB_FLAG_PI_FROM_STATES) != 0;
    // One GPU, mma.sync / FFMA tiles: nothing sits between backward() and optimizer.step(), so every job that finishes a
    // gradient (dW tiles, bias column sums, the loss reduction's db2) applies Adam (+ Polyak) in its own epilogue and the two
    // Adam phases disappear (18 -> 16 phases).  In-place parameter updates must not race with a dX GEMM that reads the
    // same weights in the same phase, so the actor's dW jobs run one phase after the dX through that layer.
    const bool fuse_adam = e->world <= 1 && !e->peer_mode && e->tf32x3 < 2;
    auto with_adam = [&](Job j, bool critic) {
        if (!fuse_adam) return j;
        const long long off = j.C - (critic ? GC : GA);
        j.adam = critic ? 2 : 1;
        j.out0 = (critic ? e->mc : e->ma) + off; j.out1 = (critic ? e->vc : e->va) + off;
        j.out2 = (critic ? PC : PA) + off; j.out3 = critic ? PT + off : nullptr;
        return j; };

    // ---- job factories -------------------------------------------------------------------------------------------
    auto actor_l0 = [&](int dyn, int b_static, bool from_states = false) {
        Job j = from_states ? b.gemm(G_FWD, E_BIAS_RELU, e->xs, ldx, 0, PA + L.actor_w0, O, 0, e->ah1, H1, 0, B, H1, O, 1, dyn)
                            : b.gemm(G_FWD, E_BIAS_RELU, e->s2buf, lds, 0, PA + L.actor_w0, O, 0, e->ah1, H1, 0, B, H1, O, 1, dyn);
        j.bias = PA + L.actor_b0; j.b_static = b_static; if (fuse_adam && !dyn && !from_states) j.stg = 1; return j; };
    auto actor_l1 = [&](int dyn) {
        Job j = b.gemm(G_FWD, E_BIAS_RELU, e->ah1, H1, 0, PA + L.actor_w1, H1, 0, e->ah2, H2, 0, B, H2, H1, 1, dyn);
        j.bias = PA + L.actor_b1; j.b_static = 1;
        if (fuse_heads) {        // post-product: heads partials [tile][2A][B] = h2 tile . [W_mu ; W_ls] slice
            j.i1 = 2 * A; j.aux1 = PA + L.actor_w_mu; j.i2 = H2; j.i3 = 0; j.out1 = e->hpart; j.i0 = B;
        }
        return j; };
    auto actor_heads = [&](int dyn) {
        Job j = b.gemm(G_FWD, E_BIAS, e->ah2, H2, 0, PA + L.actor_w_mu, H2, 0, e->heads, ldh, 0, B, 2 * A, H2, 1, dyn);
        j.bias = PA + L.actor_b_mu; j.b_static = 1; return j; };
    auto row_heads = [&](int dyn, int passes = 0) {
        Job j = Builder::zero();
        j.type = J_ROW_HEADS; j.A = e->heads; j.C = e->heads; j.lda = ldh; j.M = B; j.rows_dynamic = dyn;
        j.i3 = passes;                              // 0 = whatever RunArgs.pass_mask says, else ANDed with it
        j.out0 = e->x2; j.out1 = e->logp; j.out2 = e->xp; j.out3 = e->logp + B; j.ldc = ldx; j.i0 = O;
        if (fuse_heads) { j.i1 = e->tnH2; j.B = e->hpart; j.i2 = B; j.bias = PA + L.actor_b_mu; }
        j.ntiles = (B + kRowsHeads - 1) / kRowsHeads; return j; };
    // twin layer 0 as ONE [2*H1, O+A] GEMM, or (half = 0 / 1) only the q1 / q2 half of its output columns
    auto critic_l0 = [&](float* P, const float* X, float* H, int dyn, int b_static, int half = -1) {
        const int c0 = half < 0 ? 0 : half * H1, nc = half < 0 ? 2 * H1 : H1;
        Job j = b.gemm(G_FWD, E_BIAS_RELU, X, ldx, 0, P + L.critic_w0[0] + (size_t)c0 * OA, OA, 0, H + c0, 2 * H1, 0, B, nc, OA, 1, dyn);
        j.bias = P + L.critic_b0[0] + c0; j.b_static = b_static; return j; };
    // z0 .. z0+nz-1 of the twin layer 1 (with the Q-dot epilogue)
    // (row0, nrows): only that block of batch rows -- lets a GEMM be spread over the idle CTAs of two phases
    auto critic_l1 = [&](float* P, const float* Hin, float* Hout, float* part, int dyn, int z0, int nz, int row0 = 0, int nrows = -1) {
        if (nrows < 0) nrows = B;
        Job j = b.gemm(G_FWD, E_BIAS_RELU_DOT, Hin + (size_t)z0 * H1 + (size_t)row0 * 2 * H1, 2 * H1, H1, P + L.critic_w1[z0], H1, sW1,
                       Hout + (size_t)z0 * H2 + (size_t)row0 * 2 * H2, 2 * H2, H2, nrows, H2, H1, nz, dyn);
        j.bias = P + L.critic_b1[z0]; j.sBias = H2; j.aux0 = P + L.critic_w2[z0]; j.sAux = H2;
        j.out0 = part + (size_t)z0 * e->tnH2 * B + row0; j.i0 = B; j.b_static = 1;
        return j; };
    auto critic_dx1 = [&](bool with_dpi) {       // dh1 = (dh2 . W1) * relu'(h1)   [+ d pi partials]
        Job j = b.gemm(G_DX, E_MASK, e->dh2, 2 * H2, H2, PC + L.critic_w1[0], H1, sW1, e->dh1, 2 * H1, H1, B, H1, H2, 2);
        j.mask = e->ch1; j.ldmask = 2 * H1; j.sMask = H1; j.b_static = 1;
        if (with_dpi && fuse_dpi) {   // X[n][a] = W0cat[(z*H1 + n), O + a]: z stride H1 rows of OA floats
            j.i1 = A; j.aux1 = PC + L.critic_w0[0] + O; j.i2 = OA; j.i3 = 1; j.sAux = (long long)H1 * OA; j.out1 = e->dpart; j.i0 = B;
        }
        return j; };
    auto gather = [&](int k_offset) {
        Job j = Builder::zero();
        j.type = J_GATHER; j.M = B; j.C = e->xs; j.out0 = e->xp; j.out1 = e->x2; j.ldc = ldx; j.out2 = e->s2buf; j.ldb = lds;
        j.out3 = e->rd; j.ntiles = (B + kRowsWarp - 1) / kRowsWarp; j.k_offset = k_offset;
        if (fuse_adam) j.stg = 4;
        return j; };
    auto adam = [&](bool critic) {
        Job j = Builder::zero();
        j.type = J_ADAM;
        if (critic) { j.C = PC; j.A = GC; j.out0 = e->mc; j.out1 = e->vc; j.out2 = PT; j.K = (int)L.critic_floats; j.i0 = 1; }
        else { j.C = PA; j.A = GA; j.out0 = e->ma; j.out1 = e->va; j.out2 = nullptr; j.K = (int)L.actor_floats; j.i0 = 0; }
        {   // one tile per CTA, each walking its share of the 1024-float groups with the next group's loads in flight
            const int ng = (j.K + kThreads * 4 - 1) / (kThreads * 4), cap_t = e->grid_coop > 0 ? e->grid_coop : 148;
            j.ntiles = ng < cap_t ? ng : cap_t;
        }
        j.i2 = 1;                                  // gradient comes from the data-parallel exchange when one is active
        return j; };
    auto l2_prefetch = [&](const float* base, long long floats, int tiles) {
        Job j = Builder::zero();
        j.type = J_PREFETCH; j.A = base; j.K = (int)floats; j.ntiles = tiles;
        return j; };
    // bf16/TMA path (separate Adam phases, large nets): the CTAs idle in the last weight-gradient phase pull the optimizer
    // state the Adam phase is about to stream (one arena each: params, exp_avg, exp_avg_sq [, target]) into L2
    const bool pf_adam = e->tf32x3 == 3 && !fuse_adam && e->world <= 1;
    auto peer_signal = [&](bool critic) {
        Job j = Builder::zero();
        j.type = J_PEER_SIGNAL; j.i0 = critic ? 1 : 0; j.ntiles = 1;
        return j; };

    // actor trunk + heads + tanh-Gaussian rows; `with_critic` interleaves the online critic's forward (update_critic)
    // Tile budget of update_critic's forward part (one round = one tile per SM, `cap` tiles): the online critic's forward
    // (C.L0 / C.L1, two halves each) has slack until the backup rows, so its tiles fill the CTAs the actor / target GEMMs
    // leave idle.  When C.L0[q2] would not fit next to A.L1 and C.L1[q1] (C1: 64 + 64 + 64 > 148), C.L1[q1] and most of
    // C.L1[q2] move next to the tanh-Gaussian rows and the last row blocks of C.L1[q2] next to T.L1.
    const int cap = e->grid_coop > 0 ? e->grid_coop : 148;
    const int t_l1 = ((B + e->tile_m - 1) / e->tile_m) * e->tnH2;                 // tiles of one [B, H2] layer-1 GEMM
    const int t_l0 = ((B + e->tile_m - 1) / e->tile_m) * tnH1;                    // ... of one [B, H1] layer-0 GEMM
    const int t_rows = (B + kRowsHeads - 1) / kRowsHeads;
    int q2_late_rows = 0;                                                          // rows of C.L1[q2] deferred to the T.L1 phase
    bool spread = false;
    if (fuse_heads && t_l1 + t_l1 + t_l0 > cap && t_l1 + t_l0 <= cap) {
        spread = true;
        const int room = cap - t_rows - t_l1;                                      // tiles left next to rows_heads + C.L1[q1]
        const int blocks_fit = room > 0 ? room / e->tnH2 : 0, blocks = (B + e->tile_m - 1) / e->tile_m;
        const int late_blocks = blocks_fit >= blocks ? 0 : blocks - blocks_fit;
        q2_late_rows = late_blocks * e->tile_m;
        if (2 * t_l1 + late_blocks * e->tnH2 > cap || q2_late_rows >= B) { spread = false; q2_late_rows = 0; }
    }
    auto actor_forward = [&](int dyn, bool with_critic, int l0_static, int passes = 0, bool from_states = false) {
        if (with_critic && spread) {
            b.phase(); b.add(actor_l0(dyn, l0_static)); b.add(critic_l0(PC, e->xs, e->ch1, 0, 1, 0));
            b.phase(); b.add(actor_l1(dyn)); b.add(critic_l0(PC, e->xs, e->ch1, 0, 1, 1));
            b.phase(); b.add(row_heads(dyn, passes)); b.add(critic_l1(PC, e->ch1, e->ch2, e->qpart, 0, 0, 1));
            b.add(critic_l1(PC, e->ch1, e->ch2, e->qpart, 0, 1, 1, 0, B - q2_late_rows));
            return;
        }
        // tile budget: one round = one tile per SM.  A.L0 + C.L0[q1] (e.g. 64 + 64 tiles) now, C.L0[q2] rides as the
        // cheap second tile of the next phase (its consumer C.L1[q2] runs one phase later anyway)
        b.phase(); b.add(actor_l0(dyn, l0_static, from_states)); if (with_critic) b.add(critic_l0(PC, e->xs, e->ch1, 0, 1, 0));
        // large tiles (bf16: 128x64): A.L1 + C.L1[q1] + C.L0[q2] would need a second round (3 x 64 tiles at C2), so both
        // halves of C.L1 ride with the (8-tile) heads GEMM one phase later instead
        const int t_heads = ((B + e->tile_m - 1) / e->tile_m) * ((2 * A + e->tile_n - 1) / e->tile_n);
        const bool l1_late = with_critic && !fuse_heads && 2 * t_l1 + t_l0 > cap && t_l1 + t_l0 <= cap && t_heads + 2 * t_l1 <= cap;
        b.phase(); b.add(actor_l1(dyn));
        if (with_critic) { if (!l1_late) b.add(critic_l1(PC, e->ch1, e->ch2, e->qpart, 0, 0, 1)); b.add(critic_l0(PC, e->xs, e->ch1, 0, 1, 1)); }
        if (!fuse_heads) {
            b.phase(); b.add(actor_heads(dyn));
            if (with_critic) b.add(l1_late ? critic_l1(PC, e->ch1, e->ch2, e->qpart, 0, 0, 2) : critic_l1(PC, e->ch1, e->ch2, e->qpart, 0, 1, 1));
        }
        b.phase(); b.add(row_heads(dyn, passes));
        if (with_critic && fuse_heads) b.add(critic_l1(PC, e->ch1, e->ch2, e->qpart, 0, 1, 1));
    };

    // update_policy from "critic on (s, pi)" to Adam(actor); returns the Adam phase index
    auto policy_tail = [&](bool first_static) {
        b.phase(); b.add(critic_l0(PC, e->xp, e->ch1, 0, first_static ? 1 : 0));            // critic weights were just written by Adam
        b.phase(); b.add(critic_l1(PC, e->ch1, e->ch2, e->qpart, 0, 0, 2));
        {
            b.phase();
            Job j = Builder::zero();
            j.type = J_ROW_PI; j.M = B; j.A = e->qpart; j.i0 = B; j.i1 = e->tnH2; j.bias = PC + L.critic_b2[0];
            j.aux1 = PC + L.critic_w2[0]; j.mask = e->ch2; j.ldmask = 2 * H2; j.C = e->dh2; j.ldc = 2 * H2;
            j.out0 = e->qbuf; j.out1 = e->lparts_pi; j.aux2 = e->logp + B; j.ntiles = row_tiles;
            b.add(j);
        }
        b.phase(); b.add(critic_dx1(true));
        if (!fuse_dpi) {   // d pi = dh1 . W0[:, O:O+A]   (both critics at once: reduction over 2*H1)
            b.phase();
            // one z per critic (K = H1 each, summed by the head-gradient rows): twice the tiles of a single K = 2*H1 product
            Job j = b.gemm(G_DX, E_STORE, e->dh1, 2 * H1, H1, PC + L.critic_w0[0] + O, OA, (long long)H1 * OA, e->dpi, lddpi,
                           (long long)B * lddpi, B, A, H1, 2);
            j.b_static = 1;
            b.add(j);
        }
        {
            b.phase();
            Job j = Builder::zero();
            j.type = J_ROW_GHEAD; j.M = B; j.A = e->heads; j.lda = ldh; j.B = e->dpi; j.ldb = lddpi; j.C = e->ghead; j.ldc = ldh;
            if (fuse_dpi) { j.i1 = 2 * tnH1; j.B = e->dpart; j.i2 = B; }
            else j.sB = (long long)B * lddpi;                 // d pi = the two critics' products
            j.ntiles = (B + kRowsGhead - 1) / kRowsGhead;
            b.add(j);
        }
        // A.dh2 | A.dh1 | (nothing): the dX chain;  A.dWh | A.dW1 | A.dW0: weight gradients.  Unfused, layer L's dW rides
        // with the dX through layer L (3 phases + Adam).  Fused Adam updates the weights in place, so dW_L moves one phase
        // later than the dX that reads W_L: [dh2] [dWh, dh1] [dW1, dW0] -- still 3 phases, no Adam phase.
        auto dWh = [&]() {
            { Job j = with_adam(b.gemm(G_DW, E_STORE, e->ghead, ldh, 0, e->ah2, H2, 0, GA + L.actor_w_mu, H2, 0, 2 * A, H2, B, 1), false); j.b_static = 1; b.add(j); }
            b.add(with_adam(b.colsum(e->ghead, ldh, 0, nullptr, 0, GA + L.actor_b_mu, 0, B, 2 * A, 1), false)); };
        auto dh2 = [&]() {
            Job j = b.gemm(G_DX, E_MASK, e->ghead, ldh, 0, PA + L.actor_w_mu, H2, 0, e->adh2, H2, 0, B, H2, 2 * A, 1);
            j.mask = e->ah2; j.ldmask = H2; j.b_static = 1;
            b.add(j); };
        auto dW1 = [&]() {
            { Job j = with_adam(b.gemm(G_DW, E_STORE, e->adh2, H2, 0, e->ah1, H1, 0, GA + L.actor_w1, H1, 0, H2, H1, B, 1), false); j.b_static = 1; b.add(j); }
            b.add(with_adam(b.colsum(e->adh2, H2, 0, nullptr, 0, GA + L.actor_b1, 0, B, H2, 1), false)); };
        auto dh1 = [&]() {
            Job j = b.gemm(G_DX, E_MASK, e->adh2, H2, 0, PA + L.actor_w1, H1, 0, e->adh1, H1, 0, B, H1, H2, 1);
            j.mask = e->ah1; j.ldmask = H1; j.b_static = 1;
            b.add(j); };
        auto dW0 = [&]() {
            Job j = pi_states ? with_adam(b.gemm(G_DW, E_STORE, e->adh1, H1, 0, e->xs, ldx, 0, GA + L.actor_w0, O, 0, H1, O, B, 1), false)
                              : with_adam(b.gemm(G_DW, E_STORE, e->adh1, H1, 0, e->s2buf, lds, 0, GA + L.actor_w0, O, 0, H1, O, B, 1), false);
            if (fuse_adam && !pi_states) j.stg = 2;   // the next batch's s' is being gathered into the other buffer during this phase
            j.b_static = 1;
            b.add(j);
            b.add(with_adam(b.colsum(e->adh1, H1, 0, nullptr, 0, GA + L.actor_b0, 0, B, H1, 1), false)); };
        auto loss_pi = [&]() {
            Job j = Builder::zero();
            j.type = J_REDUCE_PI; j.A = e->lparts_pi; j.i0 = row_tiles; j.out0 = e->loss_scalar + 1; j.ntiles = 1;
            b.add(j); };
        if (fuse_adam) {
            b.phase(); dh2(); loss_pi();
            b.phase(); dWh(); dh1();
            int last = b.phase(); dW1(); dW0();
            return last;
        }
        // Unfused: the dX chain first (dh2: K = 2A, dh1: K = H2 -- short reductions), then ALL weight gradients (reductions over
        // the batch, K = B: the long tiles) in one phase.  Riding dW_L with the dX through layer L instead made every one of
        // the three phases as long as a K = B tile (C2, 128x64 tiles: 11.6 + 10.4 + 10.7 us -> 3.5 + 6.5 + 11 us).
        b.phase(); dh2(); loss_pi();
        b.phase(); dh1();
        b.phase(); dWh(); dW1(); dW0();
        if (pf_adam && e->ma && e->va) {
            b.add(l2_prefetch(PA, L.actor_floats, 8)); b.add(l2_prefetch(e->ma, L.actor_floats, 8)); b.add(l2_prefetch(e->va, L.actor_floats, 8));
        }
        int adam_phase = b.phase();
        if (e->peer_mode && actor_sync) b.add(peer_signal(false));
        { Job j = adam(false); if (!actor_sync) j.i2 = 0; b.add(j); }
        return adam_phase;
    };

    // ---------------- full update: sample/gather -> update_critic -> update_policy -> update_targets ----------------
    e->p_update.begin = (int)e->phases.size();
    {   // P0 gather (ReplayBuffer.sample, buffer/replay_buffer.py:45-54).  Runs for the first update of a launch only:
        // every later batch is prefetched by the last phase of the previous update (the indices of the whole burst
        // are known up front), so the HBM latency of the random-row gather is off the critical path.
        int p0 = b.phase();
        e->phases[p0].pad = pi_states ? 0 : 1;
        b.add(gather(0));
    }
    e->ph_crit_begin = (int)e->phases.size();
    actor_forward(0, true, /*actor W0 was just written by Adam(actor) of the previous update*/ 0, pi_states ? 1 : 0);
    b.phase(); b.add(critic_l0(PT, e->x2, e->th1, 0, 1));                                   // target critic on (s', a2)
    b.phase(); b.add(critic_l1(PT, e->th1, e->th2, e->qtpart, 0, 0, 2));
    if (spread && q2_late_rows > 0) b.add(critic_l1(PC, e->ch1, e->ch2, e->qpart, 0, 1, 1, B - q2_late_rows, q2_late_rows));
    {   // backup + dq + dh2 (sac/algorithm.py:63-72)
        b.phase();
        Job j = Builder::zero();
        j.type = J_ROW_BACKUP; j.M = B; j.A = e->qpart; j.B = e->qtpart; j.i0 = B; j.i1 = e->tnH2;
        j.bias = PC + L.critic_b2[0]; j.aux0 = PT + L.critic_b2[0]; j.aux1 = PC + L.critic_w2[0];
        j.mask = e->ch2; j.ldmask = 2 * H2; j.C = e->dh2; j.ldc = 2 * H2; j.out0 = e->qbuf; j.out1 = e->lparts_q;
        j.out2 = e->logp; j.aux2 = e->rd; j.aux3 = e->rd + B; j.ntiles = row_tiles;
        b.add(j);
    }
    {   // dh1 first (layer-0 gradients need it); the light jobs (bias / head-weight gradients as 64-wide column sums,
        // loss reduction) fit on the SMs the GEMM tiles leave idle: 128 + 8 + 8 + 1 = 145 tiles at C1
        b.phase();
        b.add(critic_dx1(false));
        const int cw_light = e->tf32x3 == 3 ? 128 : 64;       // bf16: 128 GEMM tiles of 128x64 already; 8 + 8 + 1 more still fit one round
        b.add(with_adam(b.colsum(e->dh2, 2 * H2, 0, nullptr, 0, GC + L.critic_b1[0], 0, B, 2 * H2, 1, cw_light), true));
        b.add(with_adam(b.colsum(e->ch2, 2 * H2, H2, e->qbuf + 5 * (size_t)B, B, GC + L.critic_w2[0], H2, B, H2, 2, cw_light), true));
        Job j = Builder::zero();
        j.type = J_REDUCE_Q; j.A = e->lparts_q; j.i0 = row_tiles; j.out0 = e->loss_scalar; j.out1 = GC + L.critic_b2[0];
        j.out2 = GC + L.critic_b2[1]; j.ntiles = 1;
        if (fuse_adam) {
            j.adam = 2; j.aux0 = PC + L.critic_b2[0]; j.aux1 = e->mc + L.critic_b2[0]; j.aux2 = e->vc + L.critic_b2[0];
            j.aux3 = PT + L.critic_b2[0]; j.sAux = (long long)(L.critic_b2[1] - L.critic_b2[0]);
        }
        b.add(j);
    }
    {   // weight gradients of both layers: 128 + 16 + 4 = 148 tiles at C1
        b.phase();
        // B operands of weight-gradient GEMMs are forward activations nobody writes during the backward pass: prefetchable
        auto st = [](Job j) { j.b_static = 1; return j; };
        b.add(st(with_adam(b.gemm(G_DW, E_STORE, e->dh2, 2 * H2, H2, e->ch1, 2 * H1, H1, GC + L.critic_w1[0], H1, sW1, H2, H1, B, 2), true)));
        b.add(st(with_adam(b.gemm(G_DW, E_STORE, e->dh1, 2 * H1, 0, e->xs, ldx, 0, GC + L.critic_w0[0], OA, 0, 2 * H1, OA, B, 1), true)));
        b.add(with_adam(b.colsum(e->dh1, 2 * H1, 0, nullptr, 0, GC + L.critic_b0[0], 0, B, 2 * H1, 1, 128), true));
        if (pf_adam && e->mc && e->vc) {
            b.add(l2_prefetch(PC, L.critic_floats, 4)); b.add(l2_prefetch(e->mc, L.critic_floats, 4));
            b.add(l2_prefetch(e->vc, L.critic_floats, 4)); b.add(l2_prefetch(PT, L.critic_floats, 4));
        }
    }
    if (fuse_adam) {
        e->ph_crit_adam = (int)e->phases.size() - 1;        // "the phase that ends update_critic": the fused dW + Adam phase
    } else {
        e->ph_crit_adam = (int)e->phases.size();
        b.phase(); if (e->peer_mode) b.add(peer_signal(true)); b.add(adam(true));            // Adam(critic) [+ Polyak]
    }
    e->ph_pi_begin = (int)e->phases.size();
    if (pi_states) actor_forward(0, false, 1, 2, true);    // pi = actor(states): second trunk evaluation, actor-loss pass only
    e->ph_pi_adam = policy_tail(false);
    if (!pi_states) b.add(gather(1));                      // prefetch the next update's batch (staging tiles are idle here)
    e->p_update.end = (int)e->phases.size();

    // ---------------- standalone update_policy (actor trunk recomputed) ----------------
    e->p_policy.begin = (int)e->phases.size();
    actor_forward(0, false, 0, 0, pi_states);
    policy_tail(true);
    e->p_policy.end = (int)e->phases.size();

    // ---------------- update_targets ----------------
    e->p_targets.begin = (int)e->phases.size();
    {
        b.phase();
        Job j = Builder::zero();
        j.type = J_POLYAK; j.C = PT; j.A = PC; j.K = (int)L.critic_floats;
        j.ntiles = (int)((L.critic_floats + kThreads * 4 - 1) / (kThreads * 4));
        b.add(j);
    }
    e->p_targets.end = (int)e->phases.size();

    // ---------------- forward-only programs (dynamic row count) ----------------
    e->p_actor_fwd.begin = (int)e->phases.size();
    actor_forward(1, false, 0);
    e->p_actor_fwd.end = (int)e->phases.size();
    for (int w = 0; w < 2; ++w) {
        float* P = w ? PT : PC;
        e->p_critic_fwd[w].begin = (int)e->phases.size();
        b.phase(); b.add(critic_l0(P, e->xs, e->ch1, 1, 0));
        { b.phase(); Job j = critic_l1(P, e->ch1, e->ch2, e->qpart, 1, 0, 2); j.b_static = 0; b.add(j); }
        {
            b.phase();
            Job j = Builder::zero();
            j.type = J_ROW_QSUM; j.M = B; j.rows_dynamic = 1; j.A = e->qpart; j.i0 = B; j.i1 = e->tnH2;
            j.bias = P + L.critic_b2[0]; j.out0 = e->qfw; j.ntiles = (B + kRowsWarp - 1) / kRowsWarp;
            b.add(j);
        }
        e->p_critic_fwd[w].end = (int)e->phases.size();
    }

    // ---------------- the FOLDED update program (sacb_fold.cuh): 8 phases per update, 10 with a gradient exchange ----------
    //
    //   F0*  gather(k)                                  F0'* C.L1 twin on [s|a] (layer 0 generated)      (* first update of a launch)
    //   F1   A.L1 (layer 0 on s' generated; + heads as post-product)
    //   F2   T.L1 twin (a2 / logp rows + layer 0 on [s'|a2] generated; + Q-dot)
    //   F3   C.dh1 twin (backup / dq / dh2 rows generated; column sums for db0, db1, dW2 as per-row-block partials)
    //   F4   loss_q, db2 || bias / W2 gradients from the partials || C.dW1 twin || C.dW0          [+ Adam fused, Polyak]
    //   F4a  [exchange] Adam(critic) + Polyak                                                      (world > 1 only)
    //   F5   Cpi.L1 twin (pi / logp rows + layer 0 on [s|pi] generated; + Q-dot)  ||  gather(k+1)
    //   F6   Cpi.dh1 twin (min-Q rows generated; + d pi as post-product)
    //   F7   A.dh1 (head-gradient rows + dh2 generated)  ||  loss_pi  ||  C.L1 twin of update k+1, first row blocks
    //   F8   A.dWh || A.dW1 || A.dW0 || bias gradients from the partials  [+ Adam fused]  ||  C.L1 of update k+1, remaining rows
    //   F8a  [exchange] Adam(actor)                                                                (world > 1 only)
    //
    // The online critic's forward on (s, a) does not depend on the actor step, so for update k+1 it runs on the CTAs that the
    // actor's backward GEMMs of update k leave idle (the batch is gathered during F5; every staging tile is parity
    // double-buffered).  Conditions: mma.sync / FFMA tiles, two noise draws sharing the trunk (no PI_FROM_STATES), heads and
    // d pi narrow enough for the post-products, and every reduction resident in the 4 stage buffers (H1, H2 <= 256).
    e->fold_ok = false;
    {
        const char* fe = getenv("SACB_FOLD");
        const bool want_fold = !(fe && atoi(fe) == 0);
        if (want_fold && e->tf32x3 < 2 && !pi_states && fuse_heads && fuse_dpi && A <= 8 && H1 <= kStages * kBK && H2 <= kStages * kBK &&
            H1 <= kThreads && H2 <= kThreads && OA <= kGenMaxK0 && kThreads == 256 && e->tnH2 <= 8 && 2 * tnH1 <= 16) {
            const int RB = e->row_blocks;
            auto gen_job = [&](Job j, int gen, int gsrc, int gwrite, int row0 = 0) {
                j.gen = gen; j.gsrc = gsrc; j.gwrite = gwrite; j.grow0 = row0; return j; };
            auto gather_f = [&](int k_offset, int rows_per_warp) {
                Job j = Builder::zero();
                j.type = J_GATHER; j.M = B; j.C = e->xs; j.out0 = nullptr; j.out1 = nullptr; j.ldc = ldx; j.out2 = e->s2buf; j.ldb = lds;
                j.out3 = e->rd; j.i0 = rows_per_warp; j.k_offset = k_offset; j.stg = 4 | 8;
                j.ntiles = (B + kRowsWarp * rows_per_warp - 1) / (kRowsWarp * rows_per_warp);
                return j; };
            // online critic forward on [s|a] of update k + k_offset, rows [row0, row0 + nrows)
            auto cl1_q = [&](int k_offset, int row0, int nrows) {
                Job j = b.gemm(G_FWD, E_BIAS_RELU_DOT, nullptr, 2 * H1, H1, PC + L.critic_w1[0], H1, sW1,
                               e->ch2 + (size_t)row0 * 2 * H2, 2 * H2, H2, nrows, H2, H1, 2);
                j.bias = PC + L.critic_b1[0]; j.sBias = H2; j.aux0 = PC + L.critic_w2[0]; j.sAux = H2;
                j.out0 = e->qpart + row0; j.i0 = B; j.b_static = 1; j.k_offset = k_offset;
                return gen_job(j, GEN_L0, 3, 1, row0); };
            auto reduce_q = [&]() {
                Job j = Builder::zero();
                j.type = J_REDUCE_Q; j.A = e->lparts_q; j.i0 = RB; j.out0 = e->loss_scalar; j.out1 = GC + L.critic_b2[0];
                j.out2 = GC + L.critic_b2[1]; j.ntiles = 1;
                if (fuse_adam) {
                    j.adam = 2; j.aux0 = PC + L.critic_b2[0]; j.aux1 = e->mc + L.critic_b2[0]; j.aux2 = e->vc + L.critic_b2[0];
                    j.aux3 = PT + L.critic_b2[0]; j.sAux = (long long)(L.critic_b2[1] - L.critic_b2[0]);
                }
                return j; };
            auto st = [](Job j) { j.b_static = 1; return j; };
            const int cw = 512;                                        // columns per tile of the partial-sum reductions

            e->p_fold.begin = (int)e->phases.size();
            { int p0 = b.phase(); e->phases[p0].pad = 1; b.add(gather_f(0, 1)); }
            { int p0 = b.phase(); e->phases[p0].pad = 1; b.add(cl1_q(0, 0, B)); }
            {   // F1
                b.phase();
                Job j = b.gemm(G_FWD, E_BIAS_RELU, nullptr, H1, 0, PA + L.actor_w1, H1, 0, e->ah2, H2, 0, B, H2, H1, 1);
                j.bias = PA + L.actor_b1; j.b_static = 0;              // W1 was stepped by the previous phase (F8 of update k-1)
                j.i1 = 2 * A; j.aux1 = PA + L.actor_w_mu; j.i2 = H2; j.i3 = 0; j.out1 = e->hpart; j.i0 = B;
                b.add(gen_job(j, GEN_L0, 0, 1));
            }
            {   // F2
                b.phase();
                Job j = b.gemm(G_FWD, E_BIAS_RELU_DOT, nullptr, 2 * H1, H1, PT + L.critic_w1[0], H1, sW1, e->th2, 2 * H2, H2, B, H2, H1, 2);
                j.bias = PT + L.critic_b1[0]; j.sBias = H2; j.aux0 = PT + L.critic_w2[0]; j.sAux = H2;
                j.out0 = e->qtpart; j.i0 = B; j.b_static = 1;
                b.add(gen_job(j, GEN_L0, 1, 0));
            }
            {   // F3
                b.phase();
                Job j = b.gemm(G_DX, E_MASK, e->ch2, 2 * H2, H2, PC + L.critic_w1[0], H1, sW1, e->dh1, 2 * H1, H1, B, H1, H2, 2);
                j.mask = e->ch1; j.ldmask = 2 * H1; j.sMask = H1; j.b_static = 1;
                j.colp = e->colp_b0; j.ldcolp = 2 * H1;
                b.add(gen_job(j, GEN_BACKUP, 0, 0));
            }
            {   // F4: the light jobs come first, so that the CTAs that get a second tile (148 < tiles) pair a light tile with a dW0 tile
                b.phase();
                b.add(reduce_q());
                b.add(with_adam(b.colsum(e->colp_b1, 2 * H2, 0, nullptr, 0, GC + L.critic_b1[0], 0, RB, 2 * H2, 1, cw), true));
                b.add(with_adam(b.colsum(e->colp_w2, 2 * H2, 0, nullptr, 0, GC + L.critic_w2[0], 0, RB, 2 * H2, 1, cw), true));
                b.add(with_adam(b.colsum(e->colp_b0, 2 * H1, 0, nullptr, 0, GC + L.critic_b0[0], 0, RB, 2 * H1, 1, cw), true));
                b.add(st(with_adam(b.gemm(G_DW, E_STORE, e->dh2, 2 * H2, H2, e->ch1, 2 * H1, H1, GC + L.critic_w1[0], H1, sW1, H2, H1, B, 2), true)));
                Job j = st(with_adam(b.gemm(G_DW, E_STORE, e->dh1, 2 * H1, 0, e->xs, ldx, 0, GC + L.critic_w0[0], OA, 0, 2 * H1, OA, B, 1), true));
                j.stg = 16;                                              // xs of this update's parity
                b.add(j);
            }
            if (!fuse_adam) { b.phase(); if (e->peer_mode) b.add(peer_signal(true)); b.add(adam(true)); }
            {   // F5
                b.phase();
                Job j = b.gemm(G_FWD, E_BIAS_RELU_DOT, nullptr, 2 * H1, H1, PC + L.critic_w1[0], H1, sW1, e->cp2, 2 * H2, H2, B, H2, H1, 2);
                j.bias = PC + L.critic_b1[0]; j.sBias = H2; j.aux0 = PC + L.critic_w2[0]; j.sAux = H2;
                j.out0 = e->qppart; j.i0 = B; j.b_static = 0;          // critic weights were just stepped
                b.add(gen_job(j, GEN_L0, 2, 1));
                b.add(gather_f(1, 2));
            }
            {   // F6
                b.phase();
                Job j = b.gemm(G_DX, E_MASK, e->cp2, 2 * H2, H2, PC + L.critic_w1[0], H1, sW1, e->dh1, 2 * H1, H1, B, H1, H2, 2);
                j.mask = e->cp1; j.ldmask = 2 * H1; j.sMask = H1; j.b_static = 1;
                j.i1 = A; j.aux1 = PC + L.critic_w0[0] + O; j.i2 = OA; j.i3 = 1; j.sAux = (long long)H1 * OA; j.out1 = e->dpart; j.i0 = B;
                b.add(gen_job(j, GEN_PI, 0, 0));
            }
            // rows of next update's C.L1 that fit beside A.dh1 in F7 (one round = one tile per SM); the rest rides in F8
            const int t_adh1 = RB * tnH1;
            int blocks7 = (cap - t_adh1 - 1) / (2 * e->tnH2);
            if (blocks7 < 0) blocks7 = 0;
            if (blocks7 > RB) blocks7 = RB;
            const int rows7 = blocks7 * kBM < B ? blocks7 * kBM : B;
            {   // F7
                b.phase();
                Job j = b.gemm(G_DX, E_MASK, e->ah2, H2, 0, PA + L.actor_w1, H1, 0, e->adh1, H1, 0, B, H1, H2, 1);
                j.mask = e->ah1; j.ldmask = H1; j.b_static = 1;
                j.colp = e->colp_ab0; j.ldcolp = H1;
                b.add(gen_job(j, GEN_GHEAD, 0, 0));
                Job r = Builder::zero();
                r.type = J_REDUCE_PI; r.A = e->lparts_pi; r.i0 = RB; r.out0 = e->loss_scalar + 1; r.ntiles = 1;
                b.add(r);
                if (rows7 > 0) b.add(cl1_q(1, 0, rows7));
            }
            int last = 0;
            {   // F8
                last = b.phase();
                b.add(st(with_adam(b.gemm(G_DW, E_STORE, e->adh2, H2, 0, e->ah1, H1, 0, GA + L.actor_w1, H1, 0, H2, H1, B, 1), false)));
                b.add(st(with_adam(b.gemm(G_DW, E_STORE, e->ghead, ldh, 0, e->ah2, H2, 0, GA + L.actor_w_mu, H2, 0, 2 * A, H2, B, 1), false)));
                Job j = st(with_adam(b.gemm(G_DW, E_STORE, e->adh1, H1, 0, e->s2buf, lds, 0, GA + L.actor_w0, O, 0, H1, O, B, 1), false));
                j.stg = 2;
                b.add(j);
                b.add(with_adam(b.colsum(e->colp_ab1, H2, 0, nullptr, 0, GA + L.actor_b1, 0, RB, H2, 1, cw), false));
                b.add(with_adam(b.colsum(e->colp_ab0, H1, 0, nullptr, 0, GA + L.actor_b0, 0, RB, H1, 1, cw), false));
                b.add(with_adam(b.colsum(e->colp_bh, 16, 0, nullptr, 0, GA + L.actor_b_mu, 0, RB, 2 * A, 1, 32), false));
                if (rows7 < B) b.add(cl1_q(1, rows7, B - rows7));
            }
            if (!fuse_adam) {
                last = b.phase();
                if (e->peer_mode && actor_sync) b.add(peer_signal(false));
                Job j = adam(false); if (!actor_sync) j.i2 = 0; b.add(j);
            }
            (void)last;
            e->p_fold.end = (int)e->phases.size();
            e->fold_ok = true;
        }
    }
}
