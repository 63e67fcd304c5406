// sacb_tma_host.inl -- host side of the TMA-fed bf16 GEMM tiles (csrc/sacb_tma.cuh): lays out the bf16 shadow arena,
// resolves every GEMM operand of the phase programs to its shadow, encodes one CUtensorMap per (job, operand, z).
// Included by sacb_engine.cu after the programs are defined.

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tma_encoder() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

struct RefreshItem { const float* src; unsigned short* dst; long long n; int ld, cols, ld16, pad; };
constexpr int kMaxRefresh = 16;
struct RefreshTable { RefreshItem it[kMaxRefresh]; };

// fp32 -> bf16 (round to nearest even) of up to kMaxRefresh matrices in one launch: blockIdx.y = matrix
__global__ void shadow_refresh_kernel(const RefreshTable* tab) {
    const RefreshItem it = tab->it[blockIdx.y];
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < it.n; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / it.cols;
        const int c = (int)(i - r * it.cols);
        it.dst[r * it.ld16 + c] = bf16_rn(__ldg(it.src + r * it.ld + c));
    }
}

static void refresh_shadows(sacb_engine* e, cudaStream_t st) {
    if (e->n_refresh > 0) shadow_refresh_kernel<<<dim3(64, e->n_refresh), 256, 0, st>>>((const RefreshTable*)e->d_refresh);
}

static void tma_off(sacb_engine* e) {
    e->bf_tma = false;
    e->tmaps.clear();
    e->n_refresh = 0;
    memset(&e->sh, 0, sizeof e->sh);
    for (Job& j : e->jobs) { j.tmA = j.tmB = j.c16 = j.ldc16 = j.sC16 = j.nc32 = j.mask16 = j.ldmask16 = 0; }
}

static int build_tma(sacb_engine* e, cudaStream_t st) {
    tma_off(e);
    if (e->tf32x3 != 3 || !e->pa || !e->pc || !e->pt) return 0;
    if (const char* v = getenv("SACB_TMA")) if (atoi(v) == 0) return 0;
    EncodeTiledFn enc = tma_encoder();
    if (!enc) return 0;
    const bool verbose = getenv("SACB_TMA_VERBOSE") != nullptr;
    const sacb_config& c = e->cfg;
    const sacb_layout_info& L = e->lay;
    const long long O = c.obs_dim, A = c.act_dim, H1 = c.hidden1, H2 = c.hidden2, B = c.batch_size, OA = O + A;
    auto r8 = [](long long x) { return (x + 7) & ~7LL; };

    // ---- shadow arena layout ----
    long long off = 0;
    e->shreg.clear();
    auto reg = [&](const float* base, long long rows, int ld, long long cols) {
        sacb_engine::ShReg r{base, rows, ld, (int)cols, off, (int)r8(cols)};
        off += rows * r.ld16;
        e->shreg.push_back(r);
        return r; };
    RefreshTable tab;
    memset(&tab, 0, sizeof tab);
    int nref = 0;
    auto refresh = [&](const sacb_engine::ShReg& r) {
        if (nref < kMaxRefresh) tab.it[nref++] = RefreshItem{r.base, nullptr /* patched below */, r.rows * r.cols, r.ld, r.cols, r.ld16, (int)r.off16}; };
    // NOTE: RefreshItem.pad temporarily carries off16 (< 2^31 elements); dst is patched once the arena is allocated
    Sh16 sh;
    memset(&sh, 0, sizeof sh);
    {
        auto xs = reg(e->xs, B, e->ldx, OA), xp = reg(e->xp, B, e->ldx, OA), x2 = reg(e->x2, B, e->ldx, OA), s2 = reg(e->s2buf, B, e->lds, O);
        sh.xs = (int)xs.off16; sh.xp = (int)xp.off16; sh.x2 = (int)x2.off16; sh.s2 = (int)s2.off16; sh.ldx = xs.ld16; sh.lds = s2.ld16;
        refresh(xs); refresh(xp); refresh(x2); refresh(s2);
    }
    reg(e->ah1, B, (int)H1, H1); reg(e->ah2, B, (int)H2, H2);
    reg(e->ch1, B, (int)(2 * H1), 2 * H1); reg(e->th1, B, (int)(2 * H1), 2 * H1);
    reg(e->dh2, B, (int)(2 * H2), 2 * H2); reg(e->dh1, B, (int)(2 * H1), 2 * H1);
    reg(e->ghead, B, e->ldh, 2 * A); reg(e->adh2, B, (int)H2, H2); reg(e->adh1, B, (int)H1, H1);
    auto arena = [&](float* P, bool critic, int& base16, ShMat* mats, int& nmat) {
        base16 = (int)off;
        int n = 0;
        auto mat = [&](long long woff, long long rows, long long cols) {
            auto r = reg(P + woff, rows, (int)cols, cols);
            refresh(r);
            mats[n++] = ShMat{(int)woff, (int)(rows * cols), (int)cols, (int)(r.off16 - base16), r.ld16}; };
        if (critic) { mat(L.critic_w0[0], 2 * H1, OA); mat(L.critic_w1[0], H2, H1); mat(L.critic_w1[1], H2, H1); }
        else { mat(L.actor_w0, H1, O); mat(L.actor_w1, H2, H1); mat(L.actor_w_mu, 2 * A, H2); }
        nmat = n; };
    ShMat scratch[4]; int nscratch = 0;
    arena(e->pa, false, sh.pa, sh.mat_a, sh.nmat_a);
    arena(e->pc, true, sh.pc, sh.mat_c, sh.nmat_c);
    arena(e->pt, true, sh.pt, scratch, nscratch);
    if (off >= (1LL << 31)) return 0;                       // offsets are ints on the device
    if ((size_t)off > e->sh16_elems) {
        if (e->sh16) CU(cudaFree(e->sh16));
        e->sh16 = nullptr; e->sh16_elems = 0;
        CU(cudaMalloc(&e->sh16, (size_t)off * sizeof(unsigned short)));
        e->sh16_elems = (size_t)off;
    }
    CU(cudaMemsetAsync(e->sh16, 0, (size_t)off * sizeof(unsigned short), st));
    sh.base = e->sh16;
    for (int i = 0; i < nref; ++i) { tab.it[i].dst = e->sh16 + tab.it[i].pad; tab.it[i].pad = 0; }

    // ---- operands -> shadows -> tensor maps ----
    struct Res { bool ok; bool inside; long long off16; int ld16; };      // inside && !ok: shadowed buffer, but not 16-byte aligned there
    auto resolve = [&](const float* p) {
        for (const auto& r : e->shreg) {
            if (p >= r.base && p < r.base + r.rows * r.ld) {
                const long long d = p - r.base, row = d / r.ld, col = d % r.ld;
                const long long o = r.off16 + row * r.ld16 + col;
                return Res{col < r.cols && (o & 7) == 0, true, o, r.ld16};
            }
        }
        return Res{false, false, 0, 0}; };
    const char* why = nullptr;
    auto make_map = [&](const float* p, bool rc, long long out_dim, long long red_dim, int box_out) {
        Res r = resolve(p);
        if (!r.ok) { why = "operand has no 16-byte-aligned shadow"; return 0; }
        CUtensorMap tm;
        cuuint64_t dims[2] = {(cuuint64_t)(rc ? red_dim : out_dim), (cuuint64_t)(rc ? out_dim : red_dim)};
        cuuint64_t strides[1] = {(cuuint64_t)r.ld16 * 2};
        cuuint32_t box[2] = {64u, (cuuint32_t)(rc ? box_out : 64)};
        cuuint32_t estr[2] = {1u, 1u};
        CUresult cr = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, e->sh16 + r.off16, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (cr != CUDA_SUCCESS) { why = "cuTensorMapEncodeTiled failed"; return 0; }
        e->tmaps.push_back(tm);
        return (int)e->tmaps.size(); };
    bool ok = true;
    for (Job& j : e->jobs) {
        if (j.type != J_GEMM) continue;
        const bool a_rc = j.mode != G_DW, b_rc = j.mode == G_FWD;
        // maps of one operand are adjacent: the device addresses them as tmA + z / tmB + z
        for (int z = 0; z < j.nz && ok; ++z) {
            const int ia = make_map(j.A + (size_t)z * j.sA, a_rc, j.M, j.K, kBfM);
            if (!ia) ok = false;
            else if (z == 0) j.tmA = ia;
        }
        for (int z = 0; z < j.nz && ok; ++z) {
            const int ib = make_map(j.B + (size_t)z * j.sB, b_rc, j.N, j.K, kBfN);
            if (!ib) ok = false;
            else if (z == 0) j.tmB = ib;
        }
        if (!ok) break;
    }
    if (ok) {
        for (Job& j : e->jobs) {
            if (j.type == J_GEMM || j.type == J_ROW_BACKUP || j.type == J_ROW_PI || j.type == J_ROW_GHEAD) {
                if (!j.C) continue;
                Res r = resolve(j.C);
                if (!r.inside) continue;                    // nobody reads this output through the TMA
                if (!r.ok) { ok = false; why = "GEMM output inside a shadowed buffer at an unaligned column"; break; }
                j.c16 = (int)r.off16 + 1; j.ldc16 = r.ld16;
                if (j.type == J_GEMM && j.nz == 2) {
                    Res r1 = resolve(j.C + j.sC);
                    if (!r1.ok) { ok = false; why = "twin output without shadow"; break; }
                    j.sC16 = (int)(r1.off16 - r.off16);
                }
            }
        }
    }
    if (ok) {
        // Activations that are only GEMM operands / relu masks afterwards keep no fp32 copy in this mode (the working set of
        // one update must stay L2-resident: ncu showed 71 MB of DRAM write-backs per update with every fp32 copy kept);
        // relu masks are read from the bf16 shadow.  The target critic's layer-1 output is consumed by its Q-dot epilogue only.
        auto in_buf = [&](const float* p, const float* base, long long n) { return p >= base && p < base + n; };
        bool masks16 = true;                                  // every relu mask can be read from a shadow
        for (Job& j : e->jobs) {
            if (j.type != J_GEMM || j.epi != E_MASK || !j.mask) continue;
            Res r = resolve(j.mask);
            if (r.inside && r.ok && (j.sMask & 7) == 0) { j.mask16 = (int)r.off16 + 1; j.ldmask16 = r.ld16; }
            else masks16 = false;
        }
        for (Job& j : e->jobs) {
            if (j.type != J_GEMM) continue;
            if (masks16 && j.c16 && (in_buf(j.C, e->ah1, B * H1) || in_buf(j.C, e->ah2, B * H2) || in_buf(j.C, e->ch1, B * 2 * H1) || in_buf(j.C, e->th1, B * 2 * H1)))
                j.nc32 = 1;
            if (in_buf(j.C, e->th2, B * 2 * H2)) j.nc32 = 1;
        }
    }
    if (!ok) {
        if (verbose) fprintf(stderr, "sacb: TMA operand path off (%s); precision 3 uses the staged cp.async tiles\n", why ? why : "?");
        tma_off(e);
        return 0;
    }
    if (e->tmaps.size() > e->d_tmaps_cap) {
        if (e->d_tmaps) CU(cudaFree(e->d_tmaps));
        e->d_tmaps = nullptr; e->d_tmaps_cap = 0;
        CU(cudaMalloc(&e->d_tmaps, (e->tmaps.size() + 16) * sizeof(CUtensorMap)));
        e->d_tmaps_cap = e->tmaps.size() + 16;
    }
    if (!e->d_refresh) CU(cudaMalloc(&e->d_refresh, sizeof(RefreshTable)));
    CU(cudaStreamSynchronize(st));
    CU(cudaMemcpy(e->d_tmaps, e->tmaps.data(), e->tmaps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(e->d_refresh, &tab, sizeof tab, cudaMemcpyHostToDevice));
    sh.tmaps = e->d_tmaps;
    { const char* v = getenv("SACB_TMA_PREB"); sh.pre_b = !(v && atoi(v) == 0); }
    e->sh = sh;
    e->n_refresh = nref;
    e->bf_tma = true;
    if (verbose) fprintf(stderr, "sacb: TMA operand path on: %zu tensor maps, %lld shadow elements\n", e->tmaps.size(), off);
    return 0;
}
