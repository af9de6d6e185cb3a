// Warp-specialised thread-per-element action kernel `volume_action_ws` (n_basis <= 5) with its shared-memory metric ring:
// DESIGN.md §4.2. This file is textually included by operators.cu INSIDE `namespace cb200 { namespace {` (it uses PlanDev,
// Tables, bulk_prefetch_l2 defined there); it holds device code only, the launch plumbing stays in operators.cu.
// Reference semantics replaced: source/StiffnessMatrix.cpp:83-184, source/MassMatrix.cpp:137-211 (one CTA per element,
// atomicAdd scatter) and the composition of examples/Helmholtz.hpp:28-56.
#pragma once
        __device__ __forceinline__ void cp_async8(void * smem, const void * g)
        {
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(g) : "memory");
        }
        __device__ __forceinline__ void cp_async_wait_all()
        {
            asm volatile("cp.async.commit_group;\ncp.async.wait_all;" ::: "memory");
        }

        // metric layout of the thread-per-element kernel below: [pair of values][element], rows padded to an even count
        template <int NB, int NQ, bool STIFF>
        struct TpeCfg
        {
            static constexpr int NKI = STIFF ? 3 : 1;
            static constexpr int KR = (NKI * NQ + 1) & ~1; // metric values per quadrature row, padded to even
            static constexpr int NK2 = NQ * KR / 2;        // 16-byte pairs per element
        };

        // ------------------------------------------------------------------------------------------
        // Warp-specialised persistent thread-per-element kernel (the default for n_basis <= 5).
        //
        // ncu on the lane-per-row kernel above (profiles/r01_notes.md) pins its plateau on the L1TEX / shared-memory pipe
        // (two transposes per element through a per-warp scratch, partial warps, bank conflicts: ~69 wavefronts per element)
        // and on instruction issue. Here ONE THREAD owns ONE ELEMENT: U (nb x nb) and the result live in registers, the
        // sum-factorised contractions are DFMA chains over register operands and uniform-register table values (rolled outer
        // quadrature loop + an opaque zero in the inner table index so that ptxas streams the tables through LDCU instead of
        // hoisting 60 doubles into registers): no transposes, no scratch, all lanes busy, half the instructions. A plain
        // one-CTA-per-patch version of that idea was SLOWER (0.65 ms): with ~200 registers per thread there are 8 warps per SM
        // and only 28 % of their samples were in the contractions, the rest in the latency-bound staging / assembly phases.
        // So the phases are given to different warps of a persistent CTA (256 threads, 2 CTAs per SM):
        //   * warpgroup 1 (compute, setmaxnreg 208): one element per thread. Waits for its patch buffer, pulls U[k][e] into
        //     registers, runs the contractions (metric values straight from global memory, one quadrature row ahead,
        //     L2 hits because the block was bulk-prefetched), writes the results of the element-boundary nodes back INTO THE
        //     SAME BUFFER and those of the nodes strictly inside the element (one contributor by construction) straight to y.
        //   * warpgroup 0 (helper, setmaxnreg 48): for the next patch, gathers x through the node-major global index map
        //     with 8-byte cp.async straight into that patch's buffer (no registers, no stall), prefetches index lists and
        //     metric blocks into L2; for the previous patch, runs the deterministic assembly of the element-boundary DOFs out
        //     of its buffer (fixed-width records of up to four contributions in the plan's CSR order) and writes y / the
        //     partial slots.
        // Three patch buffers rotate: filling (i+1), computing (i), assembling (i-1). Hand-offs are named barriers
        // (bar.arrive / bar.sync over the 256 threads); the helper warpgroup frees a buffer by its own program order.
        // Summation order per DOF is the plan's CSR order, exactly as in the other kernels: bitwise reproducible.
        // ------------------------------------------------------------------------------------------
        __device__ __forceinline__ void named_sync(const int id, const int n)
        {
            asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
        }
        __device__ __forceinline__ void named_arrive(const int id, const int n)
        {
            asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory");
        }
        __device__ __forceinline__ void cluster_sync_all()
        {
            asm volatile("barrier.cluster.arrive.relaxed.aligned;\nbarrier.cluster.wait.aligned;" ::: "memory");
        }
        // split form: a thread arrives at the top of patch iteration k and waits for that arrival of everybody only at the top of
        // iteration k + 1, so the two CTAs of a pair may drift by up to one patch without anybody stalling
        __device__ __forceinline__ void cluster_arrive()
        {
            asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
        }
        __device__ __forceinline__ void cluster_wait()
        {
            asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
        }
        // 16-byte metric pairs of one quadrature row that are no longer needed after quadrature point ty
        template <int NQ, int NKI, int KR>
        __host__ __device__ constexpr int pairs_done(int ty)
        {
            return ty < 0 ? 0 : (ty + 1 >= NQ ? KR / 2 : (NKI * (ty + 1)) / 2);
        }

        // The quadrature rows of one operator ("phase") for the element of this thread. Row tx re-reads U from the patch buffer
        // (conflict-free, element-fastest) and accumulates into out[]. The metric values of two rows (even / odd) sit in g0 / g1;
        // every 16-byte pair is replaced, as soon as it has been used, by the pair of the NEXT ROW OF THE SAME PARITY: row
        // tx + 2 of this phase, or row (tx & 1) of the phase that follows (gp_next, NPN pairs per row) - two row iterations
        // ahead of its use. The scale msc of a mass phase (Helmholtz: -omega^2) is folded into the table of its last back-contraction
        // (Tables::PSrow = msc * P): one DMUL per quadrature point less than scaling the metric value.
        template <int NB, int NQ, bool STIFF, int GK, int NPN, int PE>
        __device__ __forceinline__ void contract_phase(const Tables<NB, NQ, STIFF> & tab, const double * b, double (&g0)[GK], double (&g1)[GK],
                                                       const double2 * gp, const double2 * gp_next, double (&out)[NB * NB],
                                                       const double msc, const int zero, const bool keep, const bool keep_next)
        {
            // keep / keep_next: the rows of this phase / of the following phase are read by a second CTA at about the same time
            // (other field of the fused Helmholtz apply) -> normal L2 priority; otherwise stream them with an evict-first hint
            // (one load instruction with a run-time L2 cache-hint operand)
            unsigned long long pol_keep, pol_stream;
            asm("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_keep));
            asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
            auto ldm = [&](const double2 * q, const bool k) {
                double2 v;
                asm("ld.global.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(q), "l"(k ? pol_keep : pol_stream));
                return v;
            };
            using Cfg = TpeCfg<NB, NQ, STIFF>;
            constexpr int NKI = Cfg::NKI, KR = Cfg::KR;
            constexpr int NPR = KR / 2; // pairs per row of this phase
            static_assert(GK >= KR && GK >= 2 * NPN, "metric register buffer too small");
            auto do_row = [&](const int tx, double(&g)[GK], const double2 * gnext, const int npn /* pairs per row of the target */, const bool kp) {
                const int z = tx * zero; // 0 at run time; keeps the ty-indexed table loads inside the rolled loop
                // pairs the next row needs beyond what this row frees (mass -> stiffness): those slots are idle, load at once
#pragma unroll
                for (int m = NPR; m < GK / 2; ++m)
                    if (m < npn) {
                        const double2 v = ldm(gnext + m * PE, kp);
                        g[2 * m] = v.x;
                        g[2 * m + 1] = v.y;
                    }
                double pu[NB], du[STIFF ? NB : 1];
#pragma unroll
                for (int j = 0; j < NB; ++j) {
                    double s0 = 0.0, s1 = 0.0;
#pragma unroll
                    for (int ii = 0; ii < NB; ++ii) {
                        const double u = b[(ii + NB * j) * PE];
                        s0 = fma(tab.Prow[tx][ii], u, s0);
                        if (STIFF)
                            s1 = fma(tab.Drow[tx][ii], u, s1);
                    }
                    pu[j] = s0;
                    if (STIFF)
                        du[j] = s1;
                }
                double a0[NB], a1[STIFF ? NB : 1];
#pragma unroll
                for (int q = 0; q < NB; ++q) {
                    a0[q] = 0.0;
                    if (STIFF)
                        a1[q] = 0.0;
                }
#pragma unroll
                for (int ty = 0; ty < NQ; ++ty) {
                    if (STIFF) {
                        double Dx = 0.0, Dy = 0.0;
#pragma unroll
                        for (int l = 0; l < NB; ++l) {
                            Dx = fma(tab.Prow[ty + z][l], du[l], Dx);
                            Dy = fma(tab.Drow[ty + z][l], pu[l], Dy);
                        }
                        const double A = g[NKI * ty], B = g[NKI * ty + (NKI > 1 ? 1 : 0)], C = g[NKI * ty + (NKI > 2 ? 2 : 0)];
                        const double F0 = A * Dx + B * Dy;
                        const double F1 = B * Dx + C * Dy;
#pragma unroll
                        for (int q = 0; q < NB; ++q) {
                            a0[q] = fma(tab.Prow[ty + z][q], F0, a0[q]);
                            a1[q] = fma(tab.Drow[ty + z][q], F1, a1[q]);
                        }
                    }
                    else {
                        double ppu = 0.0;
#pragma unroll
                        for (int l = 0; l < NB; ++l)
                            ppu = fma(tab.Prow[ty + z][l], pu[l], ppu);
                        const double val = g[ty] * ppu; // the phase scale msc sits in PSrow
#pragma unroll
                        for (int q = 0; q < NB; ++q)
                            a0[q] = fma(tab.Prow[ty + z][q], val, a0[q]);
                    }
#pragma unroll
                    for (int m = pairs_done<NQ, NKI, KR>(ty - 1); m < pairs_done<NQ, NKI, KR>(ty); ++m)
                        if (m < npn) {
                            const double2 v = ldm(gnext + m * PE, kp);
                            g[2 * m] = v.x;
                            g[2 * m + 1] = v.y;
                        }
                }
#pragma unroll
                for (int q = 0; q < NB; ++q)
#pragma unroll
                    for (int ii = 0; ii < NB; ++ii) {
                        if (STIFF)
                            out[ii + NB * q] = fma(tab.Drow[tx][ii], a0[q], fma(tab.Prow[tx][ii], a1[q], out[ii + NB * q]));
                        else
                            out[ii + NB * q] = fma(tab.PSrow[tx][ii], a0[q], out[ii + NB * q]);
                    }
            };
#pragma unroll 1
            for (int tx = 0; tx < NQ; tx += 2) {
                const bool in0 = tx + 2 < NQ;
                do_row(tx, g0, in0 ? gp + (tx + 2) * NPR * PE : gp_next, in0 ? NPR : NPN, in0 ? keep : keep_next);
                if (tx + 1 < NQ) {
                    const bool in1 = tx + 3 < NQ;
                    do_row(tx + 1, g1, in1 ? gp + (tx + 3) * NPR * PE : gp_next + NPN * PE, in1 ? NPR : NPN, in1 ? keep : keep_next);
                }
            }
        }

        // ---- shared-memory metric ring (TMA bulk copies + mbarriers), used by the heavy fused instances -------------------
        // The register double buffer above keeps ~one quadrature row of metric data in flight per thread (ptxas sinks the
        // reloads to the end of a row) and costs 72 registers at n_basis 5. Here the compute warpgroup feeds itself through a
        // ring of RING chunks in shared memory: a chunk = up to five 16-byte pairs of one quadrature row for all 128 elements
        // of the patch, one contiguous block of the [pair][element] layout = ONE cp.async.bulk (TMA) copy completing on an
        // mbarrier. At the start of a row every thread waits for the row's chunk(s), pulls its pairs into registers (one row:
        // 36 registers instead of 72), a 128-thread named barrier says the slots are free, and thread 0 immediately issues the
        // copies of the chunks RING positions further down the stream (across phases, fields and patches). Depth is set by
        // shared memory, not by registers or by ptxas' scheduling.
        __device__ __forceinline__ unsigned smem_u32(const void * p)
        {
            return (unsigned)__cvta_generic_to_shared(p);
        }
        __device__ __forceinline__ void mbar_init(const unsigned bar, const int count)
        {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
        }
        __device__ __forceinline__ void mbar_expect_tx(const unsigned bar, const unsigned bytes)
        {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        }
        __device__ __forceinline__ void bulk_copy_g2s(const unsigned dst, const void * src, const unsigned bytes, const unsigned bar)
        {
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                         "r"(bytes), "r"(bar)
                         : "memory");
        }
        __device__ __forceinline__ void mbar_wait(const unsigned bar, const unsigned parity)
        {
            asm volatile("{\n"
                         ".reg .pred p;\n"
                         "MBAR_WAIT_LOOP:\n"
                         "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                         "@p bra MBAR_WAIT_DONE;\n"
                         "bra MBAR_WAIT_LOOP;\n"
                         "MBAR_WAIT_DONE:\n"
                         "}" ::"r"(bar),
                         "r"(parity)
                         : "memory");
        }
        // chunks of a quadrature row with NPR pairs: NH chunks of at most CP <= 5 pairs
        __host__ __device__ constexpr int ring_nh(int npr) { return (npr + 4) / 5; }
        __host__ __device__ constexpr int ring_cp(int npr) { return (npr + ring_nh(npr) - 1) / ring_nh(npr); }

        struct RingState // consumer position in the ring (uniform across the warpgroup)
        {
            const double2 * ring; // [RING][CHUNK_PAIRS][PE]
            unsigned mbar;        // shared address of the RING mbarriers
            int slot, parity;
        };

        // one operator phase fed from the ring; `issue(slot)` (thread 0 only) starts the copy of the next chunk of the stream
        // TR: per-thread ring (see contract_mass_tring below) - every thread copies its own pairs with 16-byte cp.async, one group
        // per chunk, waits with cp.async.wait_group and refills the chunks of a row after the row's values have been consumed:
        // no mbarrier, no row barrier, no leader. `issue(slot)` is then called by every thread.
        template <int NB, int NQ, bool STIFF, int PE, int RING, int CHUNK_PAIRS, bool TR = false, class IssueFn>
        __device__ __forceinline__ void contract_phase_ring(const Tables<NB, NQ, STIFF> & tab, const double (&U)[NB * NB], RingState & rs, const int e,
                                                            double (&out)[NB * NB], const double msc, const int zero, const bool leader,
                                                            IssueFn issue)
        {
            using Cfg = TpeCfg<NB, NQ, STIFF>;
            constexpr int NKI = Cfg::NKI, KR = Cfg::KR, NPR = KR / 2;
            constexpr int NH = ring_nh(NPR), CP = ring_cp(NPR);
            static_assert(CP <= CHUNK_PAIRS && NH <= RING, "ring geometry");
            // one barrier per row (two single-chunk mass rows sharing a barrier was measured: n_basis 5 fused 0.734 -> 0.786 ms)
            constexpr int RG = 1;
            // first-index contraction of a row: needs no metric data
            auto row_first = [&](const int tx, double (&pu)[NB], double (&du)[STIFF ? NB : 1]) {
#pragma unroll
                for (int j = 0; j < NB; ++j) {
                    double s0 = 0.0, s1 = 0.0;
#pragma unroll
                    for (int ii = 0; ii < NB; ++ii) {
                        const double u = U[ii + NB * j];
                        s0 = fma(tab.Prow[tx][ii], u, s0);
                        if (STIFF)
                            s1 = fma(tab.Drow[tx][ii], u, s1);
                    }
                    pu[j] = s0;
                    if (STIFF)
                        du[j] = s1;
                }
            };
            auto row_rest = [&](const int tx, const double (&g)[KR], const double (&pu)[NB], const double (&du)[STIFF ? NB : 1]) {
                const int z = tx * zero; // 0 at run time; keeps the ty-indexed table loads inside the rolled loop
                double a0[NB], a1[STIFF ? NB : 1];
#pragma unroll
                for (int q = 0; q < NB; ++q) {
                    a0[q] = 0.0;
                    if (STIFF)
                        a1[q] = 0.0;
                }
#pragma unroll
                for (int ty = 0; ty < NQ; ++ty) {
                    if (STIFF) {
                        double Dx = 0.0, Dy = 0.0;
#pragma unroll
                        for (int l = 0; l < NB; ++l) {
                            Dx = fma(tab.Prow[ty + z][l], du[l], Dx);
                            Dy = fma(tab.Drow[ty + z][l], pu[l], Dy);
                        }
                        const double A = g[NKI * ty], B = g[NKI * ty + (NKI > 1 ? 1 : 0)], C = g[NKI * ty + (NKI > 2 ? 2 : 0)];
                        const double F0 = A * Dx + B * Dy;
                        const double F1 = B * Dx + C * Dy;
#pragma unroll
                        for (int q = 0; q < NB; ++q) {
                            a0[q] = fma(tab.Prow[ty + z][q], F0, a0[q]);
                            a1[q] = fma(tab.Drow[ty + z][q], F1, a1[q]);
                        }
                    }
                    else {
                        double ppu = 0.0;
#pragma unroll
                        for (int l = 0; l < NB; ++l)
                            ppu = fma(tab.Prow[ty + z][l], pu[l], ppu);
                        const double val = g[ty] * ppu; // the phase scale msc sits in PSrow
#pragma unroll
                        for (int q = 0; q < NB; ++q)
                            a0[q] = fma(tab.Prow[ty + z][q], val, a0[q]);
                    }
                }
#pragma unroll
                for (int q = 0; q < NB; ++q)
#pragma unroll
                    for (int ii = 0; ii < NB; ++ii) {
                        if (STIFF)
                            out[ii + NB * q] = fma(tab.Drow[tx][ii], a0[q], fma(tab.Prow[tx][ii], a1[q], out[ii + NB * q]));
                        else
                            out[ii + NB * q] = fma(tab.PSrow[tx][ii], a0[q], out[ii + NB * q]);
                    }
            };
#pragma unroll 1
            for (int tx0 = 0; tx0 < NQ; tx0 += RG) {
                double g[RG][KR];
                int freed[RG * NH];
#pragma unroll
                for (int rr = 0; rr < RG; ++rr)
                    if (rr == 0 || tx0 + rr < NQ) {
#pragma unroll
                        for (int h = 0; h < NH; ++h) {
                            if constexpr (TR) {
                                if (h == 0)
                                    asm volatile("cp.async.wait_group %0;" ::"n"(RING - NH) : "memory"); // the NH oldest chunk groups have landed
                            }
                            else
                                mbar_wait(rs.mbar + 8 * rs.slot, (unsigned)rs.parity);
                            const double2 * src = rs.ring + (size_t)rs.slot * (CHUNK_PAIRS * PE) + e;
#pragma unroll
                            for (int m = 0; m < CP; ++m)
                                if (h * CP + m < NPR) {
                                    const double2 v = src[m * PE];
                                    g[rr][2 * (h * CP + m)] = v.x;
                                    g[rr][2 * (h * CP + m) + 1] = v.y;
                                }
                            freed[rr * NH + h] = rs.slot;
                            if (++rs.slot == RING) {
                                rs.slot = 0;
                                rs.parity ^= 1;
                            }
                        }
                    }
                // the first-index contraction runs while the ring reads above are still in flight
                double pu[NB], du[STIFF ? NB : 1];
                row_first(tx0, pu, du);
                if constexpr (!TR) {
                    named_sync(8, PE); // every thread of the warpgroup holds its pairs in registers: the slots are free
                    if (leader) {
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#pragma unroll
                        for (int rr = 0; rr < RG; ++rr)
                            if (rr == 0 || tx0 + rr < NQ) {
#pragma unroll
                                for (int h = 0; h < NH; ++h)
                                    issue(freed[rr * NH + h]);
                            }
                    }
                }
                row_rest(tx0, g[0], pu, du);
                if constexpr (TR) { // the row's values have been consumed (their shared-memory reads completed): refill its chunks
#pragma unroll
                    for (int h = 0; h < NH; ++h)
                        issue(freed[h]);
                }
            }
        }

        // Stiffness phase on AFFINE elements (parallelograms - every element of Mesh2D::uniform_rect): the Jacobian is constant per
        // element, so G(tx, ty) = w_tx w_ty (gA, gB, gC) with three per-element numbers (source/StiffnessMatrix.cpp:5-38 evaluates
        // the same expression point by point). The weights are folded into the back-contraction tables (PWrow = w_q P(q,.),
        // DWrow = w_q D(q,.)), so the phase needs NO metric stream (24 instead of 24 nq^2 bytes per element), no ring slot, no
        // mbarrier and no row barrier, and not one flop more than the stored-metric formulation:
        //   F0' = gA Dx + gB Dy ; F1' = gB Dx + gC Dy ; a0 += PW(ty,.) F0' ; a1 += DW(ty,.) F1' ; out += DW(tx,.) a0 + PW(tx,.) a1
        // Same numbers as the general path up to re-association (~1e-16 relative; the 1e-12 parity tests cover both paths).
        template <int NB, int NQ>
        __device__ __forceinline__ void contract_stiff_affine(const Tables<NB, NQ, true> & tab, const double (&U)[NB * NB], const double gA,
                                                              const double gB, const double gC, double (&out)[NB * NB], const int zero)
        {
#pragma unroll 1
            for (int tx = 0; tx < NQ; ++tx) {
                const int z = tx * zero; // 0 at run time; keeps the ty-indexed table loads inside the rolled loop
                double pu[NB], du[NB];
#pragma unroll
                for (int j = 0; j < NB; ++j) {
                    double s0 = 0.0, s1 = 0.0;
#pragma unroll
                    for (int ii = 0; ii < NB; ++ii) {
                        const double u = U[ii + NB * j];
                        s0 = fma(tab.Prow[tx][ii], u, s0);
                        s1 = fma(tab.Drow[tx][ii], u, s1);
                    }
                    pu[j] = s0;
                    du[j] = s1;
                }
                double a0[NB], a1[NB];
#pragma unroll
                for (int q = 0; q < NB; ++q)
                    a0[q] = a1[q] = 0.0;
#pragma unroll
                for (int ty = 0; ty < NQ; ++ty) {
                    double Dx = 0.0, Dy = 0.0;
#pragma unroll
                    for (int l = 0; l < NB; ++l) {
                        Dx = fma(tab.Prow[ty + z][l], du[l], Dx);
                        Dy = fma(tab.Drow[ty + z][l], pu[l], Dy);
                    }
                    const double F0 = gA * Dx + gB * Dy;
                    const double F1 = gB * Dx + gC * Dy;
#pragma unroll
                    for (int q = 0; q < NB; ++q) {
                        a0[q] = fma(tab.PWrow[ty + z][q], F0, a0[q]);
                        a1[q] = fma(tab.DWrow[ty + z][q], F1, a1[q]);
                    }
                }
#pragma unroll
                for (int q = 0; q < NB; ++q)
#pragma unroll
                    for (int ii = 0; ii < NB; ++ii)
                        out[ii + NB * q] = fma(tab.DWrow[tx][ii], a0[q], fma(tab.PWrow[tx][ii], a1[q], out[ii + NB * q]));
            }
        }

        // Weighted-mass phase fed straight from global memory (fused AFFINE instances with RING == 0): with the stiffness metric gone
        // (contract_stiff_affine) the register file has room for TWO mass rows per thread (KR doubles each), so the rows can come
        // in with plain 128-bit loads one row ahead of their use - no shared-memory ring, no mbarrier, no 128-thread row barrier
        // and no refill work on a leader warp (profiles/r02_notes.md §1: 20 % of the compute warps' samples in the ring kernel).
        // g0 / g1 hold rows 0 / 1 on entry (loaded by the caller BEFORE the stiffness phase, so they have landed long before);
        // every 16-byte pair is replaced by the pair of row tx + 2 as soon as its last value has been used. The blocks were
        // bulk-prefetched into L2 by the helper one patch ahead. Same arithmetic, value for value, as contract_phase_ring.
        __device__ __forceinline__ double2 ld_metric_pair(const double2 * q)
        {
            double2 v;
            asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(q));
            return v;
        }
        template <int NB, int NQ, int PE>
        __device__ __forceinline__ void contract_mass_direct(const Tables<NB, NQ, false> & tab, const double (&U)[NB * NB],
                                                             double (&g0)[TpeCfg<NB, NQ, false>::KR], double (&g1)[TpeCfg<NB, NQ, false>::KR],
                                                             const double2 * gp, double (&out)[NB * NB], const double msc, const int zero)
        {
            constexpr int KR = TpeCfg<NB, NQ, false>::KR, NPR = KR / 2;
            auto do_row = [&](const int tx, double (&g)[KR], const double2 * gnext, const bool reload) {
                const int z = tx * zero; // 0 at run time; keeps the ty-indexed table loads inside the rolled loop
                double pu[NB];
#pragma unroll
                for (int j = 0; j < NB; ++j) {
                    double s0 = 0.0;
#pragma unroll
                    for (int ii = 0; ii < NB; ++ii)
                        s0 = fma(tab.Prow[tx][ii], U[ii + NB * j], s0);
                    pu[j] = s0;
                }
                double a0[NB];
#pragma unroll
                for (int q = 0; q < NB; ++q)
                    a0[q] = 0.0;
#pragma unroll
                for (int ty = 0; ty < NQ; ++ty) {
                    double ppu = 0.0;
#pragma unroll
                    for (int l = 0; l < NB; ++l)
                        ppu = fma(tab.Prow[ty + z][l], pu[l], ppu);
                    const double val = g[ty] * ppu; // the phase scale msc sits in PSrow
#pragma unroll
                    for (int q = 0; q < NB; ++q)
                        a0[q] = fma(tab.Prow[ty + z][q], val, a0[q]);
#pragma unroll
                    for (int m = pairs_done<NQ, 1, KR>(ty - 1); m < pairs_done<NQ, 1, KR>(ty); ++m)
                        if (reload) {
                            const double2 v = ld_metric_pair(gnext + m * PE);
                            g[2 * m] = v.x;
                            g[2 * m + 1] = v.y;
                        }
                }
#pragma unroll
                for (int q = 0; q < NB; ++q)
#pragma unroll
                    for (int ii = 0; ii < NB; ++ii)
                        out[ii + NB * q] = fma(tab.PSrow[tx][ii], a0[q], out[ii + NB * q]);
            };
#pragma unroll 1
            for (int tx = 0; tx < NQ; tx += 2) {
                do_row(tx, g0, gp + (size_t)(tx + 2) * NPR * PE, tx + 2 < NQ);
                if (tx + 1 < NQ)
                    do_row(tx + 1, g1, gp + (size_t)(tx + 3) * NPR * PE, tx + 3 < NQ);
            }
        }

        // Weighted-mass phase fed from a PER-THREAD shared-memory ring (fused AFFINE instances with RING < 0, depth -RING rows):
        // every compute thread copies the rows of ITS OWN element with 16-byte cp.async (a warp's copies are one contiguous 512-byte
        // run per pair) into its own column of the ring, -RING rows ahead of their use and across patches, and tracks completion with
        // cp.async.wait_group. A thread only ever reads what it copied itself, so there is no mbarrier, no row barrier, no leader
        // and no cross-warp coupling (the TMA ring above: 12 % of the compute warps' samples on the row barrier), and unlike the
        // register-buffered variant (contract_mass_direct: 14 % long-scoreboard samples, one row ahead) the depth is not limited by
        // the register file. Same arithmetic, value for value, as contract_phase_ring.
        __device__ __forceinline__ void cp_async16(void * smem, const void * g)
        {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(g) : "memory");
        }
        template <int N>
        __device__ __forceinline__ void cp_async_wait_group()
        {
            asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
        }
        template <int NB, int NQ, int PE, int RD, class RequestFn>
        __device__ __forceinline__ void contract_mass_tring(const Tables<NB, NQ, false> & tab, const double (&U)[NB * NB], const double2 * myring,
                                                            int & slot, double (&out)[NB * NB], const double msc, const int zero,
                                                            RequestFn request)
        {
            constexpr int KR = TpeCfg<NB, NQ, false>::KR, NPR = KR / 2;
#pragma unroll 1
            for (int tx = 0; tx < NQ; ++tx) {
                const int z = tx * zero; // 0 at run time; keeps the ty-indexed table loads inside the rolled loop
                cp_async_wait_group<RD - 1>(); // the oldest of the RD row groups in flight has landed
                double g[KR];
                const double2 * src = myring + (size_t)slot * (NPR * PE);
#pragma unroll
                for (int m = 0; m < NPR; ++m) {
                    const double2 v = src[m * PE];
                    g[2 * m] = v.x;
                    g[2 * m + 1] = v.y;
                }
                double pu[NB];
#pragma unroll
                for (int j = 0; j < NB; ++j) {
                    double s0 = 0.0;
#pragma unroll
                    for (int ii = 0; ii < NB; ++ii)
                        s0 = fma(tab.Prow[tx][ii], U[ii + NB * j], s0);
                    pu[j] = s0;
                }
                double a0[NB];
#pragma unroll
                for (int q = 0; q < NB; ++q)
                    a0[q] = 0.0;
#pragma unroll
                for (int ty = 0; ty < NQ; ++ty) {
                    double ppu = 0.0;
#pragma unroll
                    for (int l = 0; l < NB; ++l)
                        ppu = fma(tab.Prow[ty + z][l], pu[l], ppu);
                    const double val = g[ty] * ppu; // the phase scale msc sits in PSrow
#pragma unroll
                    for (int q = 0; q < NB; ++q)
                        a0[q] = fma(tab.Prow[ty + z][q], val, a0[q]);
                }
                // every value of the slot has been consumed (its loads completed long ago): refill it with the row RD further on
                request(slot);
                if (++slot == RD)
                    slot = 0;
#pragma unroll
                for (int q = 0; q < NB; ++q)
#pragma unroll
                    for (int ii = 0; ii < NB; ++ii)
                        out[ii + NB * q] = fma(tab.PSrow[tx][ii], a0[q], out[ii + NB * q]);
            }
        }

        // second-phase placeholder for the single-operator instances
        struct NoTables
        {
            double pad;
        };
        template <int NB, int NQ2>
        struct Phase2
        {
            using type = Tables<NB, (NQ2 > 0 ? NQ2 : 1), false>;
        };

        // NQ2 > 0: a weighted-mass phase (scale msc) follows the first operator on the same element data - the Helmholtz
        // composite S - omega^2 M. n_fields = 2 walks [u; v] (x, y, partial strided by field, sign[f] on the result).
        struct WsArgs
        {
            const double2 * G1;
            const double2 * G2;
            const double * Gc; // AFFINE instances: per patch (3, 128) metric constants gA, gB, gC of every element slot
            const double * x;
            double * y;
            double * partial;
            long long x_stride, y_stride, partial_stride; // per field
            double c[2];                                  // result scale per field
            double msc;                                   // scale of the second (mass) phase
            int accumulate, n_patches, n_fields, zero;
        };

        // AFFINE (stiffness instances on meshes whose elements are all parallelograms): first phase = contract_stiff_affine, the
        // ring (if any) carries the mass phase only.
        template <int NB, int NQ, bool STIFF, int NQ2, int RING, bool AFFINE = false>
        __global__ void __launch_bounds__(256, 2)
        volume_action_ws(const __grid_constant__ Tables<NB, NQ, STIFF> tab, const __grid_constant__ typename Phase2<NB, NQ2>::type tab2,
                         const PlanDev plan, const __grid_constant__ WsArgs args)
        {
            using Cfg = TpeCfg<NB, NQ, STIFF>;
            using Cfg2 = TpeCfg<NB, (NQ2 > 0 ? NQ2 : 1), false>;
            constexpr int PE = 128;
            constexpr int NB2 = NB * NB;
            constexpr int NPR1 = Cfg::KR / 2, NPR2 = NQ2 > 0 ? Cfg2::KR / 2 : 0;
            constexpr int GK = 2 * (NPR1 > NPR2 ? NPR1 : NPR2); // metric registers per row buffer (RING == 0)
            constexpr int BUF = NB2 * PE; // doubles per patch buffer
            constexpr size_t g_patch1 = (size_t)Cfg::NK2 * PE, g_patch2 = NQ2 > 0 ? (size_t)Cfg2::NK2 * PE : 0; // double2 per patch
            constexpr int FULL = 1, READY = 4, HELPER = 7; // named barrier ids (0 is __syncthreads)
            constexpr int NG = (NB2 + 3) / 4;                // groups of four nodes in the global index map
            constexpr int NI = (NB > 2 ? NB - 2 : 0) * (NB > 2 ? NB - 2 : 0); // element-interior nodes
            constexpr bool HEAVY = STIFF && NB >= 5 && NQ2 > 0 && RING == 0 && !AFFINE; // register-path fused instance only
            static_assert(!AFFINE || STIFF, "AFFINE is a property of the stiffness phase");
            // fused AFFINE instances feed the mass phase from the ring (RING > 0) or straight from global memory (RING == 0)

            // shared-memory metric ring (RING > 0): chunk = CHUNK_PAIRS x PE 16-byte pairs
            constexpr int CP1 = ring_cp(NPR1), CP2 = NQ2 > 0 ? ring_cp(NPR2) : 0;
            // fused instances: split cluster barrier (the CTAs of a pair may drift by one patch) for the per-thread rings, lockstep
            // for the TMA rings (measured 0.61 -> 0.81 ms with the split barrier there: the second CTA outruns the first one's L2
            // prefetches). Compile-time on purpose: a run-time choice between the two aligned barrier forms costs ptxas the
            // uniformity of the row loops (table loads fall from LDCU to LDC: 0.72 -> 1.17 ms on the stored-metric instance).
            constexpr bool CSPLIT = RING < 0;
            constexpr bool TRM = RING < 0 && AFFINE && NQ2 > 0; // per-thread ring of whole mass rows (contract_mass_tring)
            constexpr bool TRG = RING < 0 && !TRM;              // per-thread ring of chunks, any phase (contract_phase_ring<TR>)
            constexpr int CHUNK_PAIRS = TRM ? NPR2 : (CP1 > CP2 ? CP1 : CP2);
            // patch buffers in rotation: 3 (fill i+1 | compute i | assemble i-1 at the same time), or 2 for the deep-ring heavy
            // instances, where one unit of compute is long enough for the helper to assemble i-1 and THEN refill the same buffer
            constexpr int NBUF = (RING > 2 || RING < 0) ? 2 : 3;
            constexpr int RSLOTS = RING < 0 ? -RING : RING; // RING < 0: per-thread cp.async ring of -RING mass rows (contract_mass_tring)

            extern __shared__ __align__(16) unsigned char smem_raw[];
            double * bufs = reinterpret_cast<double *>(smem_raw); // [NBUF][NB2][PE]
            int * gints = reinterpret_cast<int *>(bufs + NBUF * BUF); // [NBUF][NI][PE] global DOFs of the element-interior nodes
            double2 * ring = reinterpret_cast<double2 *>(gints + NBUF * NI * PE);                        // [RING][CHUNK_PAIRS][PE]
            unsigned long long * mbars = reinterpret_cast<unsigned long long *>(ring + RSLOTS * CHUNK_PAIRS * PE); // [RING]

            const int wg = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 7), 0); // warp-uniform by construction
            const int t = threadIdx.x & 127;
            // Fields (u, v of the fused Helmholtz apply) are dealt to the two CTAs of a thread-block cluster: CTA 2c works on u
            // and CTA 2c + 1 on v of the SAME patch sequence, kept in step by one cluster barrier per patch, so the second
            // reader of a metric block finds it in L2 (or in flight) instead of ~200 KB per CTA having to survive in L2 for a
            // whole unit (ncu: 4.4 GB of DRAM reads per apply without this, against 2.2 GB algorithmic).
            const int NF = args.n_fields;
            const int f = (int)blockIdx.x % NF, cta = (int)blockIdx.x / NF;
            const int stride = (int)gridDim.x / NF;
            const int n_iter = (args.n_patches - cta + stride - 1) / stride;
            const int accumulate = args.accumulate;

            if (wg == 0) {
                // =========================== helper warpgroup ===========================
                // register split of the 2 x 32768 budget: the stiffness contraction at n_basis 5 needs every register it can get
                // (216 / 40); the lighter instances are helper-bound and prefer 208 / 48 (measured both ways)
                if constexpr (HEAVY)
                    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
                else
                    asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
                auto issue_gather = [&](const int i) {
                    const int p = cta + i * stride;
                    const double * x = args.x + f * args.x_stride;
                    double * b = bufs + (i % NBUF) * BUF + t;
                    const int4 * ig = reinterpret_cast<const int4 *>(plan.Ig) + (size_t)p * (NG * PE) + t;
                    int4 idx[NG]; // all index loads of the element in flight at once
#pragma unroll
                    for (int gq = 0; gq < NG; ++gq)
                        idx[gq] = __ldg(ig + gq * PE);
                    // hand the global DOFs of the element-interior nodes (written by the compute thread itself) over in shared memory
                    if (NI > 0) {
                        const int n_el = __ldg(&plan.hdr[p].n_elem);
                        int * gs = gints + (i % NBUF) * (NI * PE) + t;
#pragma unroll
                        for (int k = 0; k < NB2; ++k) {
                            const int ki = k % NB, kj = k / NB;
                            if (ki > 0 && ki < NB - 1 && kj > 0 && kj < NB - 1) {
                                const int4 q4 = idx[k / 4];
                                const int gi = (k % 4 == 0) ? q4.x : (k % 4 == 1) ? q4.y : (k % 4 == 2) ? q4.z : q4.w;
                                gs[((ki - 1) + (NB - 2) * (kj - 1)) * PE] = (t < n_el) ? gi : -1;
                            }
                        }
                    }
#pragma unroll
                    for (int gq = 0; gq < NG; ++gq) {
                        cp_async8(b + (4 * gq) * PE, x + idx[gq].x);
                        if (4 * gq + 1 < NB2)
                            cp_async8(b + (4 * gq + 1) * PE, x + idx[gq].y);
                        if (4 * gq + 2 < NB2)
                            cp_async8(b + (4 * gq + 2) * PE, x + idx[gq].z);
                        if (4 * gq + 3 < NB2)
                            cp_async8(b + (4 * gq + 3) * PE, x + idx[gq].w);
                    }
                    // L2 prefetch: this patch's metric blocks (read by the compute warpgroup next iteration) and the index
                    // lists of the patch after it
                    if (f == 0) {
                        const int p2 = p + stride;
                        if (t == 0) {
                            if (!AFFINE)
                                bulk_prefetch_l2(args.G1 + (size_t)p * g_patch1, g_patch1 * sizeof(double2));
                            if (NQ2 > 0)
                                bulk_prefetch_l2(args.G2 + (size_t)p * g_patch2, g_patch2 * sizeof(double2));
                        }
                        if (t == 32 && p2 < args.n_patches) {
                            bulk_prefetch_l2(plan.Ig + (size_t)p2 * (NG * PE * 4), (size_t)NG * PE * 4 * sizeof(int));
                            const PatchHdr h2 = plan.hdr[p2];
                            bulk_prefetch_l2(plan.target + (h2.pdof_begin & ~3), ((size_t)h2.n_pdof + 4) * sizeof(int));
                            bulk_prefetch_l2(plan.cent4 + (h2.pdof_begin & ~1), ((size_t)h2.n_pdof + 2) * sizeof(uint2));
                        }
                    }
                };
                // Assembly lists of the patch to be assembled NEXT, loaded one step ahead (two-buffer instances): the loads are in
                // flight while the helper waits for its gather copies, so the assembly itself starts from registers.
                constexpr int PF = NBUF < 3 ? 8 : 0; // list entries per thread held ahead (covers 1024 listed DOFs per patch)
                uint2 prec[PF > 0 ? PF : 1];
                int ptgt[PF > 0 ? PF : 1];
                int p_npdof = 0, p_nint = 0;
                auto prefetch_lists = [&](const int i) {
                    if constexpr (PF > 0) {
                        const int p = cta + i * stride;
                        const int pdof_begin = __ldg(&plan.hdr[p].pdof_begin);
                        p_npdof = __ldg(&plan.hdr[p].n_pdof);
                        p_nint = __ldg(&plan.hdr[p].n_int);
                        const uint2 * recp = plan.cent4 + pdof_begin;
                        const int * tgtp = plan.target + pdof_begin;
#pragma unroll
                        for (int a = 0; a < PF; ++a) {
                            const int d = t + a * PE;
                            prec[a] = (d < p_npdof) ? __ldg(recp + d) : make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
                            ptgt[a] = (d < p_npdof) ? __ldg(tgtp + d) : 0;
                        }
                    }
                };
                auto assemble = [&](const int i) {
                    const int p = cta + i * stride;
                    double * y = args.y + f * args.y_stride;
                    double * partial = args.partial + f * args.partial_stride;
                    const double c = args.c[f];
                    const double * su = bufs + (i % NBUF) * BUF;
                    // one listed DOF: entries added left to right = the plan's CSR order (adding to 0.0 first is exact)
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
                    int n_pdof, n_int, first = t;
                    if constexpr (PF > 0) {
                        n_pdof = p_npdof;
                        n_int = p_nint;
#pragma unroll
                        for (int a = 0; a < PF; ++a)
                            one(t + a * PE, prec[a], ptgt[a], n_int, n_pdof);
                        first = t + PF * PE;
                        if (first >= n_pdof)
                            return;
                    }
                    else {
                        n_pdof = __ldg(&plan.hdr[p].n_pdof);
                        n_int = __ldg(&plan.hdr[p].n_int);
                    }
                    const int pdof_begin = __ldg(&plan.hdr[p].pdof_begin);
                    const uint2 * recp = plan.cent4 + pdof_begin;
                    const int * tgtp = plan.target + pdof_begin;
                    constexpr int CU = 4; // DOFs in flight per thread: record and target loads of a batch are independent
                    for (int base = first; base < n_pdof; base += CU * PE) {
                        uint2 rec[CU];
                        int tgt[CU];
#pragma unroll
                        for (int a = 0; a < CU; ++a) {
                            const int d = base + a * PE;
                            rec[a] = (d < n_pdof) ? __ldg(recp + d) : make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
                            tgt[a] = (d < n_pdof) ? __ldg(tgtp + d) : 0;
                        }
#pragma unroll
                        for (int a = 0; a < CU; ++a)
                            one(base + a * PE, rec[a], tgt[a], n_int, n_pdof);
                    }
                };

                issue_gather(0);
                cp_async_wait_all();
                named_sync(HELPER, 128);
                named_arrive(FULL + 0, 256);
                for (int i = 0; i < n_iter; ++i) {
                    if (NF > 1) { // the CTAs of a pair stay within one patch of each other (split cluster barrier) or in lockstep
                        if constexpr (!CSPLIT)
                            cluster_sync_all();
                        else {
                            if (i > 0)
                                cluster_wait();
                            cluster_arrive();
                        }
                    }
                    if (NBUF >= 3 && i + 1 < n_iter)
                        issue_gather(i + 1);
                    if (i >= 1) {
                        named_sync(READY + (i - 1) % NBUF, 256);
                        assemble(i - 1);
                    }
                    if (NBUF < 3 && i + 1 < n_iter) { // the buffer just assembled is the one to refill
                        named_sync(HELPER, 128);
                        issue_gather(i + 1);
                    }
                    prefetch_lists(i); // for assemble(i) in the next iteration; overlaps the wait for the gather copies
                    cp_async_wait_all();
                    named_sync(HELPER, 128); // every helper thread is done reading buffer i-1 and its copies for i+1 landed
                    if (i + 1 < n_iter)
                        named_arrive(FULL + (i + 1) % NBUF, 256);
                }
                named_sync(READY + (n_iter - 1) % NBUF, 256);
                assemble(n_iter - 1);
                if (CSPLIT && NF > 1)
                    cluster_wait();
            }
            else {
                // =========================== compute warpgroup: one element per thread ===========================
                if constexpr (HEAVY)
                    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
                else
                    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
                const int e = t;
                double g0[RING > 0 ? 1 : GK], g1[RING > 0 ? 1 : GK];
                // ---- shared-memory metric ring: consumer state (all threads) and producer cursor (thread 0) ----
                RingState rs{ring, smem_u32(mbars), 0, 0};
                constexpr int FIRST_PH = AFFINE ? 1 : 0; // AFFINE: the stream holds the mass rows only
                int cur_i = 0, cur_ph = FIRST_PH, cur_r = 0, cur_h = 0; // next chunk of the stream to be issued: (patch iteration, phase, row, chunk)
                auto issue = [&](const int slot) {
                    if (cur_i >= n_iter)
                        return;
                    const int p = cta + cur_i * stride;
                    const int npr = cur_ph == 0 ? NPR1 : NPR2, cp = cur_ph == 0 ? CP1 : CP2;
                    const double2 * src = (cur_ph == 0 ? args.G1 + (size_t)p * g_patch1 : args.G2 + (size_t)p * g_patch2) +
                                          (size_t)(cur_r * npr + cur_h * cp) * PE;
                    const int pairs = min(cp, npr - cur_h * cp);
                    const unsigned bytes = (unsigned)(pairs * PE * sizeof(double2));
                    const unsigned bar = rs.mbar + 8 * slot;
                    mbar_expect_tx(bar, bytes);
                    bulk_copy_g2s(smem_u32(ring + (size_t)slot * (CHUNK_PAIRS * PE)), src, bytes, bar);
                    // advance the cursor: chunk -> row -> phase -> patch
                    const int nh = cur_ph == 0 ? ring_nh(NPR1) : ring_nh(NPR2 > 0 ? NPR2 : 1);
                    const int nq = cur_ph == 0 ? NQ : NQ2;
                    if (++cur_h == nh) {
                        cur_h = 0;
                        if (++cur_r == nq) {
                            cur_r = 0;
                            if (++cur_ph == (NQ2 > 0 ? 2 : 1)) {
                                cur_ph = FIRST_PH;
                                ++cur_i;
                            }
                        }
                    }
                };
                // per-thread ring of chunks (TRG): the same cursor, advanced by every thread; one cp.async group per chunk (an empty
                // group once the stream is exhausted, so that wait_group keeps counting chunks)
                unsigned trg_dst = smem_u32(ring + e);
                // the element slot as an opaque register for the requests: ptxas otherwise rebuilds it from SR_TID (an S2R round trip) at
                // every request. Only here: the same trick applied to `e` everywhere costs the default instances 1 % (register allocation).
                int e_op = e;
                if constexpr (TRG)
                    asm volatile("" : "+r"(trg_dst), "+r"(e_op));
                auto issue_tr = [&](const int slot) {
                    if (cur_i < n_iter) {
                        const int p = cta + cur_i * stride;
                        const int npr = cur_ph == 0 ? NPR1 : NPR2, cp = cur_ph == 0 ? CP1 : CP2;
                        const double2 * src = (cur_ph == 0 ? args.G1 + (size_t)p * g_patch1 : args.G2 + (size_t)p * g_patch2) +
                                              (size_t)(cur_r * npr + cur_h * cp) * PE + e_op;
                        const int pairs = min(cp, npr - cur_h * cp);
                        const unsigned dst = trg_dst + (unsigned)slot * (unsigned)(CHUNK_PAIRS * PE * sizeof(double2));
#pragma unroll
                        for (int m = 0; m < CHUNK_PAIRS; ++m)
                            if (m < pairs)
                                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (unsigned)(m * PE * sizeof(double2))),
                                             "l"(src + m * PE)
                                             : "memory");
                        const int nh = cur_ph == 0 ? ring_nh(NPR1) : ring_nh(NPR2 > 0 ? NPR2 : 1);
                        const int nq = cur_ph == 0 ? NQ : NQ2;
                        if (++cur_h == nh) {
                            cur_h = 0;
                            if (++cur_r == nq) {
                                cur_r = 0;
                                if (++cur_ph == (NQ2 > 0 ? 2 : 1)) {
                                    cur_ph = FIRST_PH;
                                    ++cur_i;
                                }
                            }
                        }
                    }
                    asm volatile("cp.async.commit_group;" ::: "memory");
                };
                if constexpr (TRG) {
#pragma unroll
                    for (int k = 0; k < RSLOTS; ++k)
                        issue_tr(k);
                }
                if constexpr (RING > 0) {
                    if (t == 0) {
                        for (int k = 0; k < RING; ++k)
                            mbar_init(rs.mbar + 8 * k, 1);
                        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
                    }
                    named_sync(8, PE);
                    if (t == 0)
                        for (int k = 0; k < RING; ++k)
                            issue(k);
                }
                else if constexpr (!AFFINE) {
                    const double2 * gpi = args.G1 + (size_t)cta * g_patch1 + e;
#pragma unroll
                    for (int m = 0; m < NPR1; ++m) {
                        const double2 v = __ldcs(gpi + m * PE);
                        g0[2 * m] = v.x;
                        g0[2 * m + 1] = v.y;
                        const double2 w = __ldcs(gpi + (NPR1 + m) * PE);
                        g1[2 * m] = w.x;
                        g1[2 * m + 1] = w.y;
                    }
                }
                // per-thread ring (RING < 0): request cursor (patch iteration, row) and consumer slot; one cp.async group per row, an
                // empty group once the stream is exhausted so that wait_group keeps counting rows
                // Source pointer, patch skip and ring address live in registers (made opaque to ptxas, which otherwise rebuilds them
                // from the constant bank and SR_TID at every request: long-scoreboard samples on the uniform datapath).
                int tr_i = 0, tr_r = 0, tr_slot = 0;
                const double2 * tr_src = TRM ? args.G2 + (size_t)cta * g_patch2 + e : nullptr; // rows of a patch are contiguous
                int tr_skip = (stride - 1) * (int)g_patch2;                                          // last row of a patch -> next patch of this CTA
                unsigned tr_dst = smem_u32(ring + e);
                if constexpr (TRM)
                    asm volatile("" : "+l"(tr_src), "+r"(tr_skip), "+r"(tr_dst));
                auto tr_request = [&](const int slot) {
                    if constexpr (TRM) {
                        if (tr_i < n_iter) {
                            const unsigned dst = tr_dst + (unsigned)slot * (unsigned)(NPR2 * PE * sizeof(double2));
#pragma unroll
                            for (int m = 0; m < NPR2; ++m)
                                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + (unsigned)(m * PE * sizeof(double2))),
                                             "l"(tr_src + m * PE)
                                             : "memory");
                            tr_src += NPR2 * PE;
                            if (++tr_r == NQ2) {
                                tr_r = 0;
                                ++tr_i;
                                tr_src += tr_skip;
                            }
                        }
                        asm volatile("cp.async.commit_group;" ::: "memory");
                    }
                };
                if constexpr (TRM) {
#pragma unroll
                    for (int k = 0; k < -RING; ++k)
                        tr_request(k);
                }
                for (int i = 0; i < n_iter; ++i) {
                    const int p = cta + i * stride;
                    double * b = bufs + (i % NBUF) * BUF + e;
                    // first phase of the next patch of this CTA (at the very end this one again: a harmless reload)
                    const int pn = (i + 1 < n_iter) ? p + stride : p;
                    if (NF > 1) {
                        if constexpr (!CSPLIT)
                            cluster_sync_all();
                        else {
                            if (i > 0)
                                cluster_wait();
                            cluster_arrive();
                        }
                    }
                    const double2 * gp1 = args.G1 + (size_t)p * g_patch1 + e;
                    const double2 * gp1_next = args.G1 + (size_t)pn * g_patch1 + e;
                    // AFFINE: metric constants of this thread's element, requested before the wait for the patch buffer
                    double gA = 0.0, gB = 0.0, gC = 0.0;
                    if constexpr (AFFINE) {
                        const double * gc = args.Gc + (size_t)p * (3 * PE) + e;
                        gA = __ldg(gc);
                        gB = __ldg(gc + PE);
                        gC = __ldg(gc + 2 * PE);
                    }
                    named_sync(FULL + i % NBUF, 256);
                    double out[NB2];
#pragma unroll
                    for (int k = 0; k < NB2; ++k)
                        out[k] = 0.0;
                    if constexpr (AFFINE) {
                        // U in registers for both phases
                        double U[NB2];
#pragma unroll
                        for (int k = 0; k < NB2; ++k)
                            U[k] = b[k * PE];
                        if constexpr (NQ2 > 0 && RING < 0) {
                            contract_stiff_affine<NB, NQ>(tab, U, gA, gB, gC, out, args.zero);
                            contract_mass_tring<NB, NQ2, PE, -RING>(tab2, U, ring + e, tr_slot, out, args.msc, args.zero, tr_request);
                        }
                        else if constexpr (NQ2 > 0 && RING == 0) {
                            // mass rows 0 and 1 of this patch: requested now, used after the stiffness phase
                            constexpr int KR2 = Cfg2::KR;
                            double m0[KR2], m1[KR2];
                            const double2 * gp2 = args.G2 + (size_t)p * g_patch2 + e;
#pragma unroll
                            for (int m = 0; m < KR2 / 2; ++m) {
                                const double2 v = ld_metric_pair(gp2 + m * PE), w = ld_metric_pair(gp2 + (KR2 / 2 + m) * PE);
                                m0[2 * m] = v.x;
                                m0[2 * m + 1] = v.y;
                                m1[2 * m] = w.x;
                                m1[2 * m + 1] = w.y;
                            }
                            contract_stiff_affine<NB, NQ>(tab, U, gA, gB, gC, out, args.zero);
                            contract_mass_direct<NB, NQ2, PE>(tab2, U, m0, m1, gp2, out, args.msc, args.zero);
                        }
                        else {
                            contract_stiff_affine<NB, NQ>(tab, U, gA, gB, gC, out, args.zero);
                            if constexpr (NQ2 > 0)
                                contract_phase_ring<NB, NQ2, false, PE, RING, CHUNK_PAIRS>(tab2, U, rs, e, out, args.msc, args.zero, t == 0, issue);
                        }
                    }
                    else if constexpr (RING > 0) {
                        // the ring leaves room in the register file: U stays in registers for all rows of both phases
                        double U[NB2];
#pragma unroll
                        for (int k = 0; k < NB2; ++k)
                            U[k] = b[k * PE];
                        contract_phase_ring<NB, NQ, STIFF, PE, RING, CHUNK_PAIRS>(tab, U, rs, e, out, 1.0, args.zero, t == 0, issue);
                        if constexpr (NQ2 > 0)
                            contract_phase_ring<NB, NQ2, false, PE, RING, CHUNK_PAIRS>(tab2, U, rs, e, out, args.msc, args.zero, t == 0, issue);
                    }
                    else if constexpr (TRG) {
                        double U[NB2];
#pragma unroll
                        for (int k = 0; k < NB2; ++k)
                            U[k] = b[k * PE];
                        if constexpr (!AFFINE)
                            contract_phase_ring<NB, NQ, STIFF, PE, RSLOTS, CHUNK_PAIRS, true>(tab, U, rs, e, out, 1.0, args.zero, false, issue_tr);
                        if constexpr (NQ2 > 0)
                            contract_phase_ring<NB, NQ2, false, PE, RSLOTS, CHUNK_PAIRS, true>(tab2, U, rs, e, out, args.msc, args.zero, false, issue_tr);
                    }
                    else if constexpr (NQ2 > 0) {
                        const double2 * gp2 = args.G2 + (size_t)p * g_patch2 + e;
                        // loads issued while working on field f fetch data of field f (phase 2) or of the next unit (phase 1)
                        const bool shared_read = NF > 1; // another CTA reads the same block at about the same time
                        contract_phase<NB, NQ, STIFF, GK, NPR2, PE>(tab, b, g0, g1, gp1, gp2, out, 1.0, args.zero, shared_read, shared_read);
                        contract_phase<NB, NQ2, false, GK, NPR1, PE>(tab2, b, g0, g1, gp2, gp1_next, out, args.msc, args.zero, shared_read, shared_read);
                    }
                    else
                        contract_phase<NB, NQ, STIFF, GK, NPR1, PE>(tab, b, g0, g1, gp1, gp1_next, out, 1.0, args.zero, false, false);

                    double * y = args.y + f * args.y_stride;
                    const double c = args.c[f];
                    int gint[NI > 0 ? NI : 1]; // global DOFs of this element's interior nodes (-1: padding slot)
#pragma unroll
                    for (int m = 0; m < NI; ++m)
                        gint[m] = gints[(i % NBUF) * (NI * PE) + m * PE + e];
#pragma unroll
                    for (int k = 0; k < NB2; ++k) {
                        const int ki = k % NB, kj = k / NB;
                        if (ki > 0 && ki < NB - 1 && kj > 0 && kj < NB - 1) {
                            // single contributor: y (+)= c * value is exact and order-free, also as a reduction
                            const int gi = gint[(ki - 1) + (NB - 2) * (kj - 1)];
                            if (gi >= 0) {
                                const double v = c * out[k];
                                if (accumulate)
                                    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(y + gi), "d"(v) : "memory");
                                else
                                    y[gi] = v;
                            }
                        }
                        else
                            b[k * PE] = out[k];
                    }
                    named_arrive(READY + i % NBUF, 256);
                }
                if (CSPLIT && NF > 1)
                    cluster_wait();
            }
        }

