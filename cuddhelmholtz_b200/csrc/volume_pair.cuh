// Warp-specialised THREAD-PAIR-per-element action kernel `volume_action_pair` for the high orders (n_basis 6-9; BASELINE
// configs[4] order): DESIGN.md §4.2b. Textually included by operators.cu inside `namespace cb200 { namespace {`, after
// volume_ws.cuh (it reuses PlanDev, named barriers, cp.async helpers, bulk_prefetch_l2).
// Reference semantics replaced: source/StiffnessMatrix.cpp:83-184, source/MassMatrix.cpp:137-211.
//
// Why a pair. One thread per element (volume_action_ws) needs U and the result in registers: 2 nb^2 doubles = 256 registers at
// n_basis 8 - it does not exist. The lane-per-row kernel (volume_action_kernel) that served these orders spends its time in
// two shared-memory transposes per element (0.33 / 0.19 of the HBM roofline at n_basis 8). Here TWO threads (lanes l and l ^ 16
// of a warp) share an element and split its SECOND tensor index: thread h owns the columns j and the quadrature columns ty of
// its half. Per quadrature row tx
//   1. first-index contraction of the own columns (U from the patch buffer, conflict-free):  pu[jj], du[jj]
//   2. the halves swap pu / du (SHFL.BFLY), each forms Dx, Dy at ITS quadrature columns from all nb columns
//   3. metric at the own quadrature columns (stored G, or three per-element constants on affine meshes)
//   4. partial back-contraction over the own quadrature columns for ALL output columns; the halves swap the partials that
//      belong to the partner and add
//   5. first-index back-contraction into the own output columns (registers).
// = 308 DFMA per row and thread at n_basis 8 / n_quad 9 against the sum-factorised minimum of 285, 32 shuffles, no transposes,
// no scratch, ~190 registers. THE MIRROR TRICK: thread 1 works in the frame reflected in the second index (column jj <-> j =
// nb-1-jj, quadrature column tt <-> ty = nq-1-tt). GLL / Gauss nodes are symmetric, P(nq-1-t, nb-1-j) = P(t, j) and
// D(nq-1-t, nb-1-j) = -D(t, j), so BOTH threads run the same instruction stream with the same (warp-uniform, constant-bank)
// table operands; the only trace of the reflection is the sign of the off-diagonal metric term B (folded into the stored
// metric layout / the per-thread constant). Odd nb / nq: the middle column belongs to thread 0, thread 1's copy is masked.
// scripts/check_pair_algebra.py replays this thread by thread on the host against the direct formulas.
//
// Around it the same machinery as volume_action_ws: persistent CTAs of 256 threads, 2 per SM; a helper warpgroup gathers x for
// the next patch (64 elements, node-major plan) with cp.async, assembles the previous patch's element-boundary DOFs in the
// plan's CSR order (bitwise reproducible) and prefetches lists and metric blocks into L2; two patch buffers rotate. Stored
// metric data is laid out [row][pair][thread] so that every load of a warp is one contiguous 512-byte run; the stiffness streams it
// through registers one quadrature row ahead (a row is ~1200 issue cycles long at this order: deeper than an L2 round trip), the
// mass operator (short rows) through a per-thread cp.async ring of three rows. The element-interior results go straight to y: their
// ids are base + (i-1) + (j-1) * stride under H1Space's first-touch numbering (Plan::interior_affine), so the helper hands over
// two ids per element. With NQ2 > 0 the weighted-mass phase follows the (affine) stiffness phase on the same element data:
// the Helmholtz composite S - omega^2 M of one field in one launch.
#pragma once
        template <int NB, int NQ, bool STIFF>
        struct PairCfg
        {
            static constexpr int JA = (NB + 1) / 2;           // columns per thread
            static constexpr int TA = (NQ + 1) / 2;           // quadrature columns per thread
            static constexpr int NKI = STIFF ? 3 : 1;
            static constexpr int KR = (NKI * TA + 1) & ~1;    // metric values per thread and quadrature row, padded to even
            static constexpr int NPR = KR / 2;                // 16-byte pairs per thread and row
            static constexpr int PE = 64;                     // elements per patch
            static constexpr int NT = 128;                    // compute threads = metric slots per pair index
        };

        template <int NB, int NQ, bool STIFF>
        struct PairTables
        {
            static constexpr int NBP = (NB + 1) & ~1;
            double Prow[NQ][NBP];               // P(q, k)
            double Drow[STIFF ? NQ : 1][NBP];   // D(q, k)
            double PWrow[STIFF ? NQ : 1][NBP];  // w_q P(q, k): back-contractions of the affine stiffness
            double DWrow[STIFF ? NQ : 1][NBP];
            double PSrow[STIFF ? 1 : NQ][NBP];  // mass: msc * P(q, k), the phase scale folded into the last back-contraction
        };

        struct PairArgs
        {
            const double2 * G; // [patch][row tx][pair][thread slot 128]
            const double2 * G2; // fused instances: the same for the mass phase
            const double * Gc; // AFFINE: per patch (3, 64) gA, gB, gC
            const double * x;
            double * y;
            double * partial;
            double c, msc;
            int accumulate, n_patches, zero;
        };

        // The quadrature rows of ONE operator for the element of this thread pair, accumulated into out[] (see the header of this file
        // for the five steps). g: metric values of the current row (stored-metric instances without ring: loaded one row ahead, the
        // last row fetches row 0 of the next patch from gp_next); RD > 0: the rows come out of the per-thread ring at ring_t.
        template <int NB, int NQ, bool STIFF, bool AFFINE, int RD, bool UREG, bool ILP2, int GN, class Tab, class RequestFn>
        __device__ __forceinline__ void pair_rows(const Tab & tab, const double * bc, const int cstep, double (&out)[NB * ((NB + 1) / 2)],
                                                  double (&g)[GN], const double2 * gp, const double2 * gp_next, const double gA,
                                                  const double gB, const double gC, const bool mirrored, const int zero,
                                                  const double2 * ring_t, int & tr_slot, RequestFn tr_request)
        {
            using Cfg = PairCfg<NB, NQ, STIFF>;
            constexpr int PE = Cfg::PE, JA = Cfg::JA, TA = Cfg::TA, NKI = Cfg::NKI, KR = Cfg::KR, NPR = Cfg::NPR;
            static_assert(GN >= (AFFINE ? 1 : KR), "metric register buffer too small");
            double U[UREG ? NB * JA : 1];
            if constexpr (UREG) {
#pragma unroll
                for (int jj = 0; jj < JA; ++jj)
#pragma unroll
                    for (int ii = 0; ii < NB; ++ii)
                        U[ii + NB * jj] = bc[jj * cstep + ii * PE];
            }
#pragma unroll 1
            for (int tx = 0; tx < NQ; ++tx) {
                const int z = tx * zero; // 0 at run time; keeps the tt-indexed table loads inside the rolled loop
                // ---- 1. first-index contraction of the own columns ----
                double pu[JA], du[STIFF ? JA : 1];
#pragma unroll
                for (int jj = 0; jj < JA; ++jj) {
                    double s0 = 0.0, s1 = 0.0;
#pragma unroll
                    for (int ii = 0; ii < NB; ++ii) {
                        double u;
                        if constexpr (UREG)
                            u = U[ii + NB * jj];
                        else
                            u = bc[jj * cstep + ii * PE];
                        s0 = fma(tab.Prow[tx][ii], u, s0);
                        if (STIFF)
                            s1 = fma(tab.Drow[tx][ii], u, s1);
                    }
                    pu[jj] = s0;
                    if (STIFF)
                        du[jj] = s1;
                }
                if (NB % 2) { // the middle column belongs to the natural-frame thread
                    pu[JA - 1] = mirrored ? 0.0 : pu[JA - 1];
                    if (STIFF)
                        du[JA - 1] = mirrored ? 0.0 : du[JA - 1];
                }
                // ---- 2. the partner's columns ----
                double puo[JA], duo[STIFF ? JA : 1];
#pragma unroll
                for (int jj = 0; jj < JA; ++jj) {
                    puo[jj] = __shfl_xor_sync(0xffffffffu, pu[jj], 16);
                    if (STIFF)
                        duo[jj] = __shfl_xor_sync(0xffffffffu, du[jj], 16);
                }
                // ---- 2-4. own quadrature columns: second-index contraction, metric, partial back-contraction ----
                double a0o[JA], a0n[JA], a1o[STIFF ? JA : 1], a1n[STIFF ? JA : 1];
#pragma unroll
                for (int q = 0; q < JA; ++q) {
                    a0o[q] = a0n[q] = 0.0;
                    if (STIFF)
                        a1o[q] = a1n[q] = 0.0;
                }
                const double2 * gnext = (tx + 1 < NQ) ? gp + (size_t)(tx + 1) * (NPR * Cfg::NT) : gp_next;
                if constexpr (RD > 0) { // this row's pairs out of the ring (the oldest of the RD groups in flight)
                    asm volatile("cp.async.wait_group %0;" ::"n"(RD - 1) : "memory");
                    const double2 * src = ring_t + (size_t)tr_slot * (NPR * Cfg::NT);
#pragma unroll
                    for (int m = 0; m < NPR; ++m) {
                        const double2 v = src[m * Cfg::NT];
                        g[2 * m] = v.x;
                        g[2 * m + 1] = v.y;
                    }
                }
                // ILP2 (affine stiffness): all second-index sums first - pu / du are dead before the back-contraction accumulators
                // come alive, which is what lets U stay in registers at n_basis 8 (for the stored-metric and mass instances the
                // same order cost ptxas the uniform table loads, LDCU -> LDC, and was slower: 0.94 -> 1.32 ms)
                double DxA[ILP2 ? TA : 1], DyA[ILP2 ? TA : 1];
                if constexpr (ILP2) {
#pragma unroll
                    for (int tt = 0; tt < TA; ++tt)
                        DxA[tt] = DyA[tt] = 0.0;
#pragma unroll
                    for (int jj = 0; jj < JA; ++jj)
#pragma unroll
                        for (int tt = 0; tt < TA; ++tt) {
                            DxA[tt] = fma(tab.Prow[tt + z][jj], du[jj], DxA[tt]);
                            DyA[tt] = fma(tab.Drow[tt + z][jj], pu[jj], DyA[tt]);
                        }
#pragma unroll
                    for (int jj = 0; jj < JA; ++jj)
#pragma unroll
                        for (int tt = 0; tt < TA; ++tt) {
                            DxA[tt] = fma(tab.Prow[tt + z][NB - 1 - jj], duo[jj], DxA[tt]);
                            DyA[tt] = fma(tab.Drow[tt + z][NB - 1 - jj], puo[jj], DyA[tt]);
                        }
                }
#pragma unroll
                for (int tt = 0; tt < TA; ++tt) {
                    const bool dead = (NQ % 2) && tt == TA - 1 && mirrored; // the middle quadrature column belongs to thread 0
                    if (STIFF) {
                        double Dx = 0.0, Dy = 0.0;
                        if constexpr (ILP2) {
                            Dx = DxA[tt];
                            Dy = DyA[tt];
                        }
                        else {
#pragma unroll
                            for (int jj = 0; jj < JA; ++jj) {
                                Dx = fma(tab.Prow[tt + z][jj], du[jj], Dx);
                                Dy = fma(tab.Drow[tt + z][jj], pu[jj], Dy);
                            }
#pragma unroll
                            for (int jj = 0; jj < JA; ++jj) {
                                Dx = fma(tab.Prow[tt + z][NB - 1 - jj], duo[jj], Dx);
                                Dy = fma(tab.Drow[tt + z][NB - 1 - jj], puo[jj], Dy);
                            }
                        }
                        double F0, F1;
                        if constexpr (AFFINE) {
                            F0 = gA * Dx + gB * Dy;
                            F1 = gB * Dx + gC * Dy;
                            if ((NQ % 2) && tt == TA - 1) {
                                F0 = dead ? 0.0 : F0;
                                F1 = dead ? 0.0 : F1;
                            }
#pragma unroll
                            for (int q = 0; q < JA; ++q) {
                                a0o[q] = fma(tab.PWrow[tt + z][q], F0, a0o[q]);
                                a0n[q] = fma(tab.PWrow[tt + z][NB - 1 - q], F0, a0n[q]);
                                a1o[q] = fma(tab.DWrow[tt + z][q], F1, a1o[q]);
                                a1n[q] = fma(tab.DWrow[tt + z][NB - 1 - q], F1, a1n[q]);
                            }
                        }
                        else {
                            const double A = g[3 * tt], B = g[3 * tt + 1], C = g[3 * tt + 2]; // B sign-flipped, dead column zeroed in the layout
                            F0 = A * Dx + B * Dy;
                            F1 = B * Dx + C * Dy;
#pragma unroll
                            for (int q = 0; q < JA; ++q) {
                                a0o[q] = fma(tab.Prow[tt + z][q], F0, a0o[q]);
                                a0n[q] = fma(tab.Prow[tt + z][NB - 1 - q], F0, a0n[q]);
                                a1o[q] = fma(tab.Drow[tt + z][q], F1, a1o[q]);
                                a1n[q] = fma(tab.Drow[tt + z][NB - 1 - q], F1, a1n[q]);
                            }
                        }
                    }
                    else {
                        double ppu = 0.0;
#pragma unroll
                        for (int jj = 0; jj < JA; ++jj)
                            ppu = fma(tab.Prow[tt + z][jj], pu[jj], ppu);
#pragma unroll
                        for (int jj = 0; jj < JA; ++jj)
                            ppu = fma(tab.Prow[tt + z][NB - 1 - jj], puo[jj], ppu);
                        const double val = g[tt] * ppu; // the phase scale sits in PSrow
#pragma unroll
                        for (int q = 0; q < JA; ++q) {
                            a0o[q] = fma(tab.Prow[tt + z][q], val, a0o[q]);
                            a0n[q] = fma(tab.Prow[tt + z][NB - 1 - q], val, a0n[q]);
                        }
                    }
                    // metric pairs that are no longer needed: fetch the same pairs of the next row (or of the next patch)
                    if constexpr (!AFFINE && RD == 0) {
#pragma unroll
                        for (int m = pairs_done<TA, NKI, KR>(tt - 1); m < pairs_done<TA, NKI, KR>(tt); ++m) {
                            const double2 v = ld_metric_pair(gnext + m * Cfg::NT);
                            g[2 * m] = v.x;
                            g[2 * m + 1] = v.y;
                        }
                    }
                }
                if constexpr (RD > 0) { // the row's values have been consumed: refill its slot with the row RD further on
                    tr_request(tr_slot);
                    if (++tr_slot == RD)
                        tr_slot = 0;
                }
                // ---- 4b. the partner's share of the own output columns ----
#pragma unroll
                for (int q = 0; q < JA; ++q) {
                    a0o[q] += __shfl_xor_sync(0xffffffffu, a0n[q], 16);
                    if (STIFF)
                        a1o[q] += __shfl_xor_sync(0xffffffffu, a1n[q], 16);
                }
                // ---- 5. first-index back-contraction into the own output columns ----
#pragma unroll
                for (int q = 0; q < JA; ++q)
#pragma unroll
                    for (int ii = 0; ii < NB; ++ii) {
                        if (STIFF) {
                            if constexpr (AFFINE)
                                out[ii + NB * q] = fma(tab.DWrow[tx][ii], a0o[q], fma(tab.PWrow[tx][ii], a1o[q], out[ii + NB * q]));
                            else
                                out[ii + NB * q] = fma(tab.Drow[tx][ii], a0o[q], fma(tab.Prow[tx][ii], a1o[q], out[ii + NB * q]));
                        }
                        else
                            out[ii + NB * q] = fma(tab.PSrow[tx][ii], a0o[q], out[ii + NB * q]);
                    }
            }

        }

        // RD > 0: the metric rows come through a PER-THREAD shared-memory ring of RD rows (16-byte cp.async of the thread's own
        // pairs, cp.async.wait_group, no barrier - as contract_mass_tring in volume_ws.cuh) instead of through registers one row
        // ahead: the rows of the mass operator are short (176 DFMA per thread at n_basis 8), one row of lead does not cover an
        // L2 / DRAM round trip (ncu: 3.8 long-scoreboard warps per issued instruction).
        // NQ2 > 0: a weighted-mass phase (n_quad NQ2, metric args.G2 through the per-thread ring of RD2 rows, scale folded into
        // tab2.PSrow) follows the stiffness phase on the same element data and accumulates into the same registers: the Helmholtz
        // composite S - omega^2 M of one field in one launch (one gather, one assembly instead of two).
        template <int NB, int NQ, bool STIFF, bool AFFINE, int RD = 0, int NQ2 = 0, int RD2 = 0>
        __global__ void __launch_bounds__(256, 2)
        volume_action_pair(const __grid_constant__ PairTables<NB, NQ, STIFF> tab, const __grid_constant__ PairTables<NB, (NQ2 > 0 ? NQ2 : 1), false> tab2,
                           const PlanDev plan, const __grid_constant__ PairArgs args)
        {
            using Cfg = PairCfg<NB, NQ, STIFF>;
            constexpr int PE = Cfg::PE, NB2 = NB * NB, JA = Cfg::JA, TA = Cfg::TA, NKI = Cfg::NKI, KR = Cfg::KR, NPR = Cfg::NPR;
            constexpr int BUF = NB2 * PE;                       // doubles per patch buffer
            constexpr int NBUF = 2;                             // assemble i-1, refill the same buffer for i+1, while i is computed
            constexpr int NG = (NB2 + 3) / 4;                   // groups of four nodes in the global index map
            constexpr int NGH = (NG + 1) / 2;                   // groups per helper half
            constexpr int NI = 2;                               // ids handed over per element: nodes (1,1) and (1,2) - the ids of the nodes strictly
                                                                // inside an element are base + (i-1) + (j-1) * stride (Plan::interior_affine)
            constexpr size_t g_patch = (size_t)NQ * NPR * Cfg::NT; // double2 per patch
            constexpr int FULL = 1, READY = 4, HELPER = 7;      // named barrier ids
            // U (own columns) in registers for all rows where the register file allows it; otherwise re-read from the patch buffer
            // every row (ncu at n_basis 8: the compute warps then sit on the MIO queue / short scoreboard - LDS + SHFL - 45 % of the time)
            // (measured at 1024^2: n_basis 6 stiffness 0.371 -> 0.355 ms; the mass kernels gain only together with the metric ring)
            constexpr bool ILP2 = STIFF && AFFINE && NB >= 7 && NB <= 8;
            constexpr bool UREG = STIFF ? (NB <= 6 || ILP2) : (NB <= 8 && RD > 0);
            constexpr bool UREG2 = NB <= 8; // second (mass) phase, always fed from the ring
            // the ONE per-thread metric ring of the kernel serves the phase that has a depth: the mass operator of a stand-alone mass
            // instance (RD) or the mass phase of a fused instance (RD2)
            using Cfg2 = PairCfg<NB, (NQ2 > 0 ? NQ2 : 1), false>;
            constexpr int RDR = RD > 0 ? RD : RD2, NQR = RD > 0 ? NQ : NQ2, NPRR = RD > 0 ? NPR : Cfg2::NPR;
            constexpr size_t g_patchR = (size_t)NQR * NPRR * Cfg::NT;
            // (a stored-metric stiffness phase in front of the mass phase was tried: its row buffer stays live across the second phase,
            // ptxas spills ~800 bytes and the tables fall to LDC at n_basis 8 - the fused form is offered on affine meshes only)
            static_assert(NQ2 == 0 || (STIFF && AFFINE && RD == 0 && RD2 > 0), "fused instances: affine stiffness phase, mass phase with ring");
            static_assert(RD2 == 0 || NQ2 > 0, "RD2 is the ring depth of the second phase");
            static_assert(!AFFINE || STIFF, "AFFINE is a property of the stiffness operator");
            static_assert(NB >= 3, "pair kernel: n_basis >= 3");

            extern __shared__ __align__(16) unsigned char smem_raw[];
            double * bufs = reinterpret_cast<double *>(smem_raw);     // [NBUF][NB2][PE]
            int * gints = reinterpret_cast<int *>(bufs + NBUF * BUF); // [NBUF][2][PE] global DOFs of the interior nodes (1,1) and (1,2) of every element
            double2 * mring = reinterpret_cast<double2 *>(gints + NBUF * NI * PE); // [RD][NPR][128] per-thread metric ring
            static_assert(RD == 0 || !AFFINE, "the metric ring serves the stored-metric instances");
            (void)tab2;
            static_assert((NBUF * NI * PE) % 4 == 0, "ring alignment");

            const int wg = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 7), 0);
            const int t = threadIdx.x & 127;
            const int cta = (int)blockIdx.x, stride = (int)gridDim.x;
            const int n_iter = (args.n_patches - cta + stride - 1) / stride;
            const int accumulate = args.accumulate;

            if (wg == 0) {
                // =========================== helper warpgroup ===========================
                asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
                const int he = t & (PE - 1); // element slot served by this helper thread; the two halves split its node groups
                auto gather_half = [&](const int i, auto half) {
                    constexpr int H = decltype(half)::value;
                    const int p = cta + i * stride;
                    double * b = bufs + (i % NBUF) * BUF + he;
                    int * gs = gints + (i % NBUF) * (NI * PE) + he;
                    const int4 * ig = reinterpret_cast<const int4 *>(plan.Ig) + (size_t)p * (NG * PE) + he;
                    const int n_el = __ldg(&plan.hdr[p].n_elem);
#pragma unroll
                    for (int c0 = 0; c0 < NGH; c0 += 4) {
                        int4 idx[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int gq = H * NGH + c0 + u;
                            if (c0 + u < NGH && gq < NG)
                                idx[u] = __ldg(ig + gq * PE);
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int gq = H * NGH + c0 + u;
                            if (c0 + u < NGH && gq < NG) {
#pragma unroll
                                for (int w = 0; w < 4; ++w) {
                                    const int k = 4 * gq + w;
                                    if (k < NB2) {
                                        const int gi = w == 0 ? idx[u].x : w == 1 ? idx[u].y : w == 2 ? idx[u].z : idx[u].w;
                                        cp_async8(b + k * PE, args.x + gi);
                                        if (k == 1 + NB)
                                            gs[0] = (he < n_el) ? gi : -1;
                                        if (k == 1 + 2 * NB)
                                            gs[PE] = gi;
                                    }
                                }
                            }
                        }
                    }
                };
                auto issue_gather = [&](const int i) {
                    if (t < PE)
                        gather_half(i, std::integral_constant<int, 0>{});
                    else
                        gather_half(i, std::integral_constant<int, 1>{});
                    const int p = cta + i * stride, p2 = p + stride;
                    if (t == 0 && !AFFINE)
                        bulk_prefetch_l2(args.G + (size_t)p * g_patch, g_patch * sizeof(double2));
                    if (t == 32 && p2 < args.n_patches) {
                        bulk_prefetch_l2(plan.Ig + (size_t)p2 * (NG * PE * 4), (size_t)NG * PE * 4 * sizeof(int));
                        const PatchHdr h2 = plan.hdr[p2];
                        bulk_prefetch_l2(plan.target + (h2.pdof_begin & ~3), ((size_t)h2.n_pdof + 4) * sizeof(int));
                        bulk_prefetch_l2(plan.cent4 + (h2.pdof_begin & ~1), ((size_t)h2.n_pdof + 2) * sizeof(uint2));
                    }
                };
                // assembly lists of the patch to be assembled next, held in registers one step ahead
                constexpr int PF = 8;
                uint2 prec[PF];
                int ptgt[PF];
                int p_npdof = 0, p_nint = 0;
                auto prefetch_lists = [&](const int i) {
                    const int p = cta + i * stride;
                    const int pdof_begin = __ldg(&plan.hdr[p].pdof_begin);
                    p_npdof = __ldg(&plan.hdr[p].n_pdof);
                    p_nint = __ldg(&plan.hdr[p].n_int);
                    const uint2 * recp = plan.cent4 + pdof_begin;
                    const int * tgtp = plan.target + pdof_begin;
#pragma unroll
                    for (int a = 0; a < PF; ++a) {
                        const int d = t + a * 128;
                        prec[a] = (d < p_npdof) ? __ldg(recp + d) : make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
                        ptgt[a] = (d < p_npdof) ? __ldg(tgtp + d) : 0;
                    }
                };
                auto assemble = [&](const int i) {
                    const int p = cta + i * stride;
                    double * y = args.y;
                    double * partial = args.partial;
                    const double c = args.c;
                    const double * su = bufs + (i % NBUF) * BUF;
                    // one listed DOF: entries added left to right = the plan's CSR order
                    auto one = [&](const int d, const uint2 rec, const int tgt, const int n_int, const int n_pdof) {
                        const unsigned c0 = rec.x & 0xFFFFu, c1 = rec.x >> 16, c2 = rec.y & 0xFFFFu, c3 = rec.y >> 16;
                        double sum = c0 != 0xFFFFu ? su[c0] : 0.0;
                        if (c1 != 0xFFFFu)
                            sum += su[c1];
                        if (c2 != 0xFFFFu)
                            sum += su[c2];
                        if (c3 < 0xFFFEu)
                            sum += su[c3];
                        else if (c3 == 0xFFFEu) { // more than four contributions (high-valence vertex): rest of the CSR row
                            const PatchHdr hdr = plan.hdr[p];
                            const uint16_t * cp = plan.cptr + hdr.cptr_begin;
                            const uint16_t * ce = plan.cent + (size_t)hdr.elem_begin * NB2;
                            for (int k = __ldg(cp + d) + 3, e = __ldg(cp + d + 1); k < e; ++k)
                                sum += su[__ldg(ce + k)];
                        }
                        if (d < n_int) {
                            const double v = c * sum;
                            y[tgt] = accumulate ? (y[tgt] + v) : v;
                        }
                        else if (d < n_pdof)
                            partial[tgt] = sum;
                    };
                    const int n_pdof = p_npdof, n_int = p_nint;
#pragma unroll
                    for (int a = 0; a < PF; ++a)
                        one(t + a * 128, prec[a], ptgt[a], n_int, n_pdof);
                    const int first = t + PF * 128;
                    if (first >= n_pdof)
                        return;
                    const int pdof_begin = __ldg(&plan.hdr[p].pdof_begin);
                    const uint2 * recp = plan.cent4 + pdof_begin;
                    const int * tgtp = plan.target + pdof_begin;
                    constexpr int CU = 4;
                    for (int base = first; base < n_pdof; base += CU * 128) {
                        uint2 rec[CU];
                        int tgt[CU];
#pragma unroll
                        for (int a = 0; a < CU; ++a) {
                            const int d = base + a * 128;
                            rec[a] = (d < n_pdof) ? __ldg(recp + d) : make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
                            tgt[a] = (d < n_pdof) ? __ldg(tgtp + d) : 0;
                        }
#pragma unroll
                        for (int a = 0; a < CU; ++a)
                            one(base + a * 128, rec[a], tgt[a], n_int, n_pdof);
                    }
                };

                issue_gather(0);
                cp_async_wait_all();
                named_sync(HELPER, 128);
                named_arrive(FULL + 0, 256);
                for (int i = 0; i < n_iter; ++i) {
                    if (i >= 1) {
                        named_sync(READY + (i - 1) % NBUF, 256);
                        assemble(i - 1);
                    }
                    if (i + 1 < n_iter) { // the buffer just assembled is the one to refill
                        named_sync(HELPER, 128);
                        issue_gather(i + 1);
                    }
                    prefetch_lists(i); // for assemble(i) in the next iteration; overlaps the wait for the gather copies
                    cp_async_wait_all();
                    named_sync(HELPER, 128);
                    if (i + 1 < n_iter)
                        named_arrive(FULL + (i + 1) % NBUF, 256);
                }
                named_sync(READY + (n_iter - 1) % NBUF, 256);
                assemble(n_iter - 1);
            }
            else {
                // =========================== compute warpgroup: one element per thread pair ===========================
                asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
                const int e = (t >> 5) * 16 + (t & 15); // element slot: 16 per warp
                const int h = (t >> 4) & 1;             // half: lanes 0-15 work in the natural frame, lanes 16-31 in the mirrored one
                const bool mirrored = h != 0;
                // node (i, own column jj) sits at b[col0 + jj * cstep + i * PE]
                const int col0 = mirrored ? (NB - 1) * NB * PE : 0;
                const int cstep = mirrored ? -NB * PE : NB * PE;
                const int zero = args.zero;

                double g[AFFINE ? 2 : KR]; // metric values of the current quadrature row (stored-metric instances)
                // per-thread ring: running source pointer (rows of a patch are contiguous), patch skip, ring address - in registers
                const double2 * ringG = RD > 0 ? args.G : args.G2;
                const double2 * tr_src = ringG + (size_t)cta * g_patchR + t;
                int tr_skip = (stride - 1) * (int)g_patchR, tr_i = 0, tr_r = 0, tr_slot = 0;
                unsigned tr_dst = (unsigned)__cvta_generic_to_shared(mring + t);
                if constexpr (RDR > 0)
                    asm volatile("" : "+l"(tr_src), "+r"(tr_skip), "+r"(tr_dst));
                auto tr_request = [&](const int slot) {
                    if (tr_i < n_iter) {
                        const unsigned dst = tr_dst + (unsigned)slot * (unsigned)(NPRR * Cfg::NT * sizeof(double2));
#pragma unroll
                        for (int m = 0; m < NPRR; ++m)
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (unsigned)(m * Cfg::NT * sizeof(double2))),
                                         "l"(tr_src + m * Cfg::NT)
                                         : "memory");
                        tr_src += NPRR * Cfg::NT;
                        if (++tr_r == NQR) {
                            tr_r = 0;
                            ++tr_i;
                            tr_src += tr_skip;
                        }
                    }
                    asm volatile("cp.async.commit_group;" ::: "memory");
                };
                if constexpr (RDR > 0) {
#pragma unroll
                    for (int k = 0; k < RDR; ++k)
                        tr_request(k);
                }
                if constexpr (!AFFINE && RD == 0) {
                    const double2 * gp0 = args.G + (size_t)cta * g_patch + t;
#pragma unroll
                    for (int m = 0; m < NPR; ++m) {
                        const double2 v = ld_metric_pair(gp0 + m * Cfg::NT);
                        g[2 * m] = v.x;
                        g[2 * m + 1] = v.y;
                    }
                }
                for (int i = 0; i < n_iter; ++i) {
                    const int p = cta + i * stride;
                    const int pn = (i + 1 < n_iter) ? p + stride : p; // at the very end: a harmless reload
                    double * b = bufs + (i % NBUF) * BUF + e;
                    const double2 * gp = args.G + (size_t)p * g_patch + t;
                    const double2 * gp_next = args.G + (size_t)pn * g_patch + t;
                    double gA = 0.0, gB = 0.0, gC = 0.0;
                    if constexpr (AFFINE) {
                        const double * gc = args.Gc + (size_t)p * (3 * PE) + e;
                        gA = __ldg(gc);
                        gB = __ldg(gc + PE);
                        gC = __ldg(gc + 2 * PE);
                        gB = mirrored ? -gB : gB; // the off-diagonal term changes sign in the mirrored frame
                    }
                    named_sync(FULL + i % NBUF, 256);
                    double out[NB * JA]; // out[ii + NB * qq]
#pragma unroll
                    for (int k = 0; k < NB * JA; ++k)
                        out[k] = 0.0;
                    const double * bc = b + col0;
                    pair_rows<NB, NQ, STIFF, AFFINE, RD, UREG, ILP2>(tab, bc, cstep, out, g, gp, gp_next, gA, gB, gC, mirrored, zero, mring + t,
                                                                   tr_slot, tr_request);
                    if constexpr (NQ2 > 0) { // weighted-mass phase of the fused composite, scale in tab2.PSrow
                        double g2[Cfg2::KR];
                        pair_rows<NB, NQ2, false, false, RD2, UREG2, false>(tab2, bc, cstep, out, g2, nullptr, nullptr, 0.0, 0.0, 0.0, mirrored,
                                                                            zero, mring + t, tr_slot, tr_request);
                    }

                    // ---- results: element-interior nodes straight to y (one contributor), the rest back into the buffer ----
                    double * y = args.y;
                    const double c = args.c;
                    // interior node (ii, j): id = base + (ii - 1) + (j - 1) * stride, with j - 1 = q - 1 (natural frame) or nb - 2 - q (mirrored)
                    const int * gi_p = gints + (i % NBUF) * (NI * PE) + e;
                    const int gi_b = gi_p[0], gi_s = gi_p[PE] - gi_b;
                    const int gi_col0 = gi_b + (mirrored ? (NB - 2) * gi_s : -gi_s), gi_cstep = mirrored ? -gi_s : gi_s; // id of (1, column q) = gi_col0 + q * gi_cstep
#pragma unroll
                    for (int q = 0; q < JA; ++q) {
                        const bool own = !((NB % 2) && q == JA - 1) || !mirrored; // odd n_basis: the middle column is thread 0's
#pragma unroll
                        for (int ii = 0; ii < NB; ++ii) {
                            if (ii > 0 && ii < NB - 1 && q > 0) {
                                const int gi = gi_col0 + q * gi_cstep + (ii - 1);
                                if (gi_b >= 0 && own) {
                                    const double v = c * out[ii + NB * q];
                                    if (accumulate)
                                        asm volatile("red.global.add.f64 [%0], %1;" ::"l"(y + gi), "d"(v) : "memory");
                                    else
                                        y[gi] = v;
                                }
                            }
                            else if (own)
                                b[col0 + q * cstep + ii * PE] = out[ii + NB * q];
                        }
                    }
                    named_arrive(READY + i % NBUF, 256);
                }
            }
        }
