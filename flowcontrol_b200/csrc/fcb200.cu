// fcb200.cu — hand-written sm_100a kernels + C ABI for the ensemble-batched
// IMEX Navier-Stokes step of FlowControl (see include/fcb200.h for the entry
// points and the reference call sites they replace).
//
// Data layout: every ensemble field is X[row * ldb + b] (dof-major, trajectory
// innermost, FP64), ldb = B rounded up to a multiple of 32; padding columns are
// kept at zero.  In every kernel a warp owns 32 consecutive trajectories of one
// row / cell / tile, so indices, geometry and matrix values are warp-uniform
// (broadcast loads) and all state traffic is 256-byte coalesced.
//
// One time step (reference: FlowSolver.step, flowsolver.py:703-799), one CUDA graph:
//   k_ctrl_add      rhs rows += sum_k u_ctrl_k (f_k - A[:,Gamma] s_k)            (the rhs itself was left in Z by the previous step)
//   k_front_sweep   forward sweep, one launch per elimination-tree level:  y_t, u_t = sum_children u_c - E_t y_t
//   k_gather_sum    right-hand side of the merged top of the tree
//   k_front_sweep   backward sweep, one launch per level: x_t = F11^-1 y_t - G_t x_struct(t); also writes x in canonical
//                   numbering (the new state) and raises the per-trajectory non-finite flag
//   k_bc_fill       Dirichlet rows of the new state
//   k_spmm_mma      Crank-Nicolson only: rhs rows = E u_n (block-sparse SpMM on the FP64 tensor cores)
//   k_element_patch per patch of cells: convection N(u) (7-point Radon rule) + mass M u accumulated in shared memory,
//                   a = (2/dt) M u - 2 N(u),  b = -(1/2dt) M u + N(u); emits the NEXT step's rhs rows a_n + b_{n-1}
//                   (Crank-Nicolson: E u_n - N(u_n)) in solver order, b_n, and the energy partials u.(M u)
//   k_patch_merge   nodes shared between patches: sum of the patches' partials
//   k_measure       sensors (sparse rows) + energy reduction (fixed order)
// Closed loop adds k_controller before and k_log (series row + running cost sums) after, inside the same graph.
// k_rhs_build forms the rhs from a and b for the first (BDF1) step after fcb_set_state(order = 1) only.

#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <climits>
#include <cstring>
#include <string>
#include <vector>

#include "fcb200.h"

namespace {

thread_local std::string g_create_error;

// ----------------------------------------------------------------------------------------------
// constant tables (filled by fcb_create): Radon 7-point rule and P2 shape functions
// ----------------------------------------------------------------------------------------------
__constant__ double c_phi[7][6];      // phi_a(q)
__constant__ double c_dphi[7][6][2];  // reference gradients
__constant__ double c_w[7];           // quadrature weights (sum = 1/2)
__constant__ double c_mass[6][6];     // reference mass matrix (int phi_a phi_b over the unit triangle)

// ----------------------------------------------------------------------------------------------
// kernels
// ----------------------------------------------------------------------------------------------
// ----------------------------------------------------------------------------------------------
// Element kernel, patch form: one CTA = one patch of ~28 neighbouring cells x 32 trajectories; each of its
// EP_WARPS warps owns a compact sub-patch (a quarter of the cells) and integrates its cells one after the
// other with NO synchronisation: the node rows (a_x, a_y, b_x, b_y) of a sub-patch are accumulated in the
// warp's own rows of shared memory, so a node shared by 6 cells costs one global write instead of 6
// read-modify-writes and no atomics or colouring are needed.  After one __syncthreads the rows of a node that
// several sub-patches touched are summed in a fixed order; nodes interior to the patch go straight to global
// memory, nodes shared with other patches go to a scratch slot and k_patch_merge sums the (2-4) partials.
// The gathers of a warp's next cell are issued before the arithmetic of the current one.
// Fused output (bprev != NULL): instead of a the kernel writes the NEXT step's right-hand side rows
// Z[row(dof)] = a + b_{n-1} (solver order), so no separate rhs pass reads a and b again.
// ----------------------------------------------------------------------------------------------
constexpr int EP_WARPS = 4;
#ifndef FCB_EP_MINCTAS
#define FCB_EP_MINCTAS 2   // CTAs per SM the element kernel is compiled for (2 => 255 registers per thread)
#endif
#ifndef FCB_EP_PREFETCH
#define FCB_EP_PREFETCH 1  // 1: the next cell's node values are prefetched into registers while the current cell is integrated
#endif

struct PatchArgs {
    const int* pcell_ptr;          // [npatch*EP_WARPS+1] cell ranges of every warp
    const int* pcnode;             // [nslots*6] P2 node ids of every cell slot (cells listed warp by warp)
    const double* pgeo;            // [nslots*5] Jinv (4) and |det J| of every cell slot
    const unsigned char* plnode;   // [nslots*6] accumulator row (inside the CTA) of the cell's nodes
    const int* pnode_ptr;          // [npatch+1] into the per-node arrays below (unique nodes of the patch)
    const int* pnode_dst;          // >= 0: node id (interior to the patch), < 0: -(scratch slot + 1)
    const unsigned char* psrc;     // [4 per node] accumulator rows to sum (255 = none), first entry always valid
    const int* prow;               // [3 per node] solver rows of its ux, uy, p dofs (< 0: none / Dirichlet)
    const int* pacc_rows;          // [npatch] accumulator rows the patch uses
    const double* u;               // [2nN, ldb]
    double* a;
    double* b;
    double* scratch;               // [nslots, 4, ldb]
    double* epart;                 // [npatch, ldb]
    const double* bprev;           // b of the previous state (fused BDF rhs), else NULL
    double* Zb;
    int tab_off;                   // byte offset of the node tables (dst, src, 3 solver rows per node) behind the accumulators
    int mode;                      // 0: store a and b; 1: Zb rows = a + bprev, store b (BDF); 2: Zb rows += a (Crank-Nicolson)
    int nN, nV, ldb;
    double ca, cb, na, nb;         // a = ca M u + na N(u), b = cb M u + nb N(u)
};

// a/b rows of one node go out: either as a and b, or (fused) as next-step rhs rows and b
// rows of one node go out.  mode 0: a and b; mode 1 (fused BDF): next-step rhs rows Zb = a (+ bprev if ADD_BPREV, else the
// caller's a already contains it) and b; mode 2 (Crank-Nicolson): Zb rows = E u_n + a (k_spmm_mma wrote E u_n there; the
// patch kernel seeds its accumulators with it, the merge kernel adds in place)
template <bool ADD_BPREV>
__device__ __forceinline__ void emit_node(int mode, const double* bprev, double* Zb, int rx, int ry, int rp, double* a, double* bout,
                                          int nN, size_t ldb, int nd, int b, double ax, double ay, double bx, double by) {
    const size_t ox = (size_t)nd * ldb + b, oy = (size_t)(nd + nN) * ldb + b;
    if (mode == 2) {
        // ADD_BPREV (k_patch_merge): add to E u_n in place; otherwise the accumulators were seeded with it
        if (rx >= 0) Zb[(size_t)rx * ldb + b] = ADD_BPREV ? Zb[(size_t)rx * ldb + b] + ax : ax;
        if (ry >= 0) Zb[(size_t)ry * ldb + b] = ADD_BPREV ? Zb[(size_t)ry * ldb + b] + ay : ay;
        if (rp >= 0) Zb[(size_t)rp * ldb + b] = 0.0;
        return;
    }
    if (mode == 1) {
        if (rx >= 0) Zb[(size_t)rx * ldb + b] = ADD_BPREV ? ax + bprev[ox] : ax;
        if (ry >= 0) Zb[(size_t)ry * ldb + b] = ADD_BPREV ? ay + bprev[oy] : ay;
        if (rp >= 0) Zb[(size_t)rp * ldb + b] = 0.0;  // continuity rows of the rhs are zero (the solve left the pressure there)
    } else {
        a[ox] = ax;
        a[oy] = ay;
    }
    bout[ox] = bx;
    bout[oy] = by;
}

struct CellIn {
    int ln[6];
    double g00, g01, g10, g11, det;
    double ux[6], uy[6];
};

// Two-deep software pipeline per warp: the node ids of cell r+2 are fetched while the node values of cell r+1 are in
// flight and cell r is being integrated; every per-cell table is stored slot by slot, so no load depends on another
// index load.
__device__ __forceinline__ void patch_load_ids(const PatchArgs& p, int slot, int (&nd)[6]) {
#pragma unroll
    for (int i = 0; i < 6; ++i) nd[i] = __ldg(p.pcnode + (size_t)slot * 6 + i);
}
__device__ __forceinline__ void patch_load_cell(const PatchArgs& p, int slot, const int (&nd)[6], int b, CellIn& c) {
    const size_t ldb = (size_t)p.ldb;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        c.ln[i] = __ldg(p.plnode + (size_t)slot * 6 + i);
        c.ux[i] = p.u[(size_t)nd[i] * ldb + b];
        c.uy[i] = p.u[(size_t)(nd[i] + p.nN) * ldb + b];
    }
    const double* g = p.pgeo + (size_t)slot * 5;
    c.g00 = __ldg(g); c.g01 = __ldg(g + 1); c.g10 = __ldg(g + 2); c.g11 = __ldg(g + 3); c.det = __ldg(g + 4);
}

// grid = (npatch, ldb/32), block = (32, EP_WARPS), dynamic smem = max accumulator rows * 4 * 32 doubles
template <bool NONLINEAR>
__global__ void __launch_bounds__(32 * EP_WARPS, FCB_EP_MINCTAS) k_element_patch(const PatchArgs p) {
    extern __shared__ __align__(16) double acc[];  // [row][4][32]
    const int lane = threadIdx.x, w = threadIdx.y;
    const int b = blockIdx.y * 32 + lane;
    const int c0 = __ldg(p.pcell_ptr + blockIdx.x * EP_WARPS + w), c1 = __ldg(p.pcell_ptr + blockIdx.x * EP_WARPS + w + 1);
    const int n0 = __ldg(p.pnode_ptr + blockIdx.x), nn = __ldg(p.pnode_ptr + blockIdx.x + 1) - n0;
    const int nrows = __ldg(p.pacc_rows + blockIdx.x);
    CellIn cur, nxt;
    int ids[6];
    // second level of dependent loads, all issued before anything waits: node ids of the first cell, the first 32
    // entries of this warp's node tables (the cp.async of the solver rows follows below)
    const int chunk = (nn + EP_WARPS - 1) / EP_WARPS;
    const int j_lo = w * chunk, j_hi = min(nn, (w + 1) * chunk);
    int l_dst0 = -1;
    unsigned l_src0 = 0xffffffffu;
    if (c0 < c1) patch_load_ids(p, c0, ids);
    int l_rx0 = -1, l_ry0 = -1;  // Crank-Nicolson: the solver rows whose E u_n values seed the accumulators
    if (j_lo + lane < j_hi) {
        l_dst0 = __ldg(p.pnode_dst + n0 + j_lo + lane);
        l_src0 = __ldg(reinterpret_cast<const unsigned*>(p.psrc) + n0 + j_lo + lane);
        if (p.mode == 2) {
            l_rx0 = __ldg(p.prow + (size_t)(n0 + j_lo + lane) * 3);
            l_ry0 = __ldg(p.prow + (size_t)(n0 + j_lo + lane) * 3 + 1);
        }
    }
    // accumulators start at zero, except (fused right-hand side) the a rows of the patch's own nodes, which start from
    // b_{n-1}: those 256-byte row segments go straight from global to shared memory (cp.async, 16 bytes per lane: lanes
    // 0-15 the x row, 16-31 the y row), all in flight together with the first cell's gathers.  (Staged through
    // registers, eight nodes at a time, this prologue was 40 % of the kernel's warp time.)
    // node tables of the patch (destination, accumulator rows, solver rows), staged for the write-out: the solver rows
    // by 4-byte cp.async, destination and sources through the registers the preload below needs anyway
    int* s_dst = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(acc) + p.tab_off);
    unsigned* s_src = reinterpret_cast<unsigned*>(s_dst + nn);
    int* s_row = s_dst + 2 * nn;
    if (p.mode != 0)
        for (int j = 3 * j_lo + lane; j < 3 * j_hi; j += 32) {
            const uint32_t sd = (uint32_t)__cvta_generic_to_shared(s_row + j);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sd), "l"(p.prow + (size_t)n0 * 3 + j) : "memory");
        }
    // third level: the first cell's values (its ids have been in flight since above), then the ids of the second cell
    if (c0 < c1) patch_load_cell(p, c0, ids, b, cur);
    if (c0 + 1 < c1) patch_load_ids(p, c0 + 1, ids);
    for (int j0 = j_lo; j0 < j_hi; j0 += 32) {
        const int cnt = min(32, j_hi - j0);
        int l_dst = l_dst0, l_rx = l_rx0, l_ry = l_ry0;
        unsigned l_src = l_src0;
        if (j0 > j_lo && lane < cnt) {
            l_dst = __ldg(p.pnode_dst + n0 + j0 + lane);
            l_src = __ldg(reinterpret_cast<const unsigned*>(p.psrc) + n0 + j0 + lane);
            if (p.mode == 2) {
                l_rx = __ldg(p.prow + (size_t)(n0 + j0 + lane) * 3);
                l_ry = __ldg(p.prow + (size_t)(n0 + j0 + lane) * 3 + 1);
            }
        }
        if (lane < cnt) {
            s_dst[j0 + lane] = l_dst;
            s_src[j0 + lane] = l_src;
        }
        if (p.mode != 0) {
            const size_t ldb = (size_t)p.ldb;
            const int hy = lane >> 4, l16 = lane & 15;
            for (int i = 0; i < cnt; ++i) {
                const int dst = __shfl_sync(0xffffffffu, l_dst, i);
                const unsigned src = __shfl_sync(0xffffffffu, l_src, i);
                // the 256-byte row segment this half-warp seeds its accumulator row with: b_{n-1} of the node (BDF) or the
                // solver row of E u_n (Crank-Nicolson); nodes finished by k_patch_merge and eliminated rows start at zero
                const double* g = nullptr;
                if (p.mode == 1) {
                    if (dst >= 0) g = p.bprev + (size_t)(dst + hy * p.nN) * ldb;
                } else {
                    const int rx = __shfl_sync(0xffffffffu, l_rx, i), ry = __shfl_sync(0xffffffffu, l_ry, i);
                    const int r = hy ? ry : rx;
                    if (dst >= 0 && r >= 0) g = p.Zb + (size_t)r * ldb;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const unsigned rk = (src >> (8 * k)) & 255u;
                    if (rk != 255u) {  // warp-uniform
                        double* t = acc + (size_t)rk * 128 + lane;
                        if (k == 0 && g) {
                            const uint32_t sdst = (uint32_t)__cvta_generic_to_shared(acc + (size_t)rk * 128 + hy * 32 + 2 * l16);
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sdst), "l"(g + blockIdx.y * 32 + 2 * l16) : "memory");
                        } else if (k == 0) {
                            acc[(size_t)rk * 128 + hy * 32 + 2 * l16] = 0.0;
                            acc[(size_t)rk * 128 + hy * 32 + 2 * l16 + 1] = 0.0;
                        } else {
                            t[0] = 0.0; t[32] = 0.0;
                        }
                        t[64] = 0.0; t[96] = 0.0;
                    }
                }
            }
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (p.mode == 0) {
        for (int i = w; i < nrows * 4; i += EP_WARPS) acc[i * 32 + lane] = 0.0;
    }
    __syncthreads();
    double e_acc = 0.0;
    for (int r = c0; r < c1; ++r) {
#if FCB_EP_PREFETCH
        if (r + 1 < c1) patch_load_cell(p, r + 1, ids, b, nxt);
        if (r + 2 < c1) patch_load_ids(p, r + 2, ids);
#else
        if (r > c0) patch_load_cell(p, r, ids, b, cur);           // no register prefetch: more CTAs per SM hide the latency
        if (r + 1 < c1) patch_load_ids(p, r + 1, ids);
#endif
        double rx[6], ry[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) { rx[i] = 0.0; ry[i] = 0.0; }
        if (NONLINEAR) {
#pragma unroll
            for (int q = 0; q < 7; ++q) {
                double vx = 0.0, vy = 0.0, ax0 = 0.0, ax1 = 0.0, ay0 = 0.0, ay1 = 0.0;
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    vx = fma(c_phi[q][i], cur.ux[i], vx);
                    vy = fma(c_phi[q][i], cur.uy[i], vy);
                    ax0 = fma(c_dphi[q][i][0], cur.ux[i], ax0);
                    ax1 = fma(c_dphi[q][i][1], cur.ux[i], ax1);
                    ay0 = fma(c_dphi[q][i][0], cur.uy[i], ay0);
                    ay1 = fma(c_dphi[q][i][1], cur.uy[i], ay1);
                }
                // physical gradients: d_j u = sum_k (d_ref_k u) Jinv[k][j]
                const double dux_dx = ax0 * cur.g00 + ax1 * cur.g10, dux_dy = ax0 * cur.g01 + ax1 * cur.g11;
                const double duy_dx = ay0 * cur.g00 + ay1 * cur.g10, duy_dy = ay0 * cur.g01 + ay1 * cur.g11;
                const double wq = c_w[q] * cur.det;
                const double cx = wq * (vx * dux_dx + vy * dux_dy);
                const double cy = wq * (vx * duy_dx + vy * duy_dy);
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    rx[i] = fma(c_phi[q][i], cx, rx[i]);
                    ry[i] = fma(c_phi[q][i], cy, ry[i]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            double mx = 0.0, my = 0.0;
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                mx = fma(c_mass[i][j], cur.ux[j], mx);
                my = fma(c_mass[i][j], cur.uy[j], my);
            }
            mx *= cur.det;
            my *= cur.det;
            e_acc = fma(cur.ux[i], mx, e_acc);
            e_acc = fma(cur.uy[i], my, e_acc);
            double* s = acc + (size_t)cur.ln[i] * 128 + lane;  // rows of this warp's sub-patch: nobody else touches them
            s[0] += p.ca * mx + p.na * rx[i];
            s[32] += p.ca * my + p.na * ry[i];
            s[64] += p.cb * mx + p.nb * rx[i];
            s[96] += p.cb * my + p.nb * ry[i];
        }
#if FCB_EP_PREFETCH
        cur = nxt;
#endif
    }
    __syncthreads();
    // write-out: every warp takes a contiguous chunk of the patch's nodes (tables staged in shared memory by the
    // prologue, so nothing here waits on global memory); the rows go out four nodes at a time
    const size_t ldb = (size_t)p.ldb;
    {
#pragma unroll 4
        for (int j = j_lo; j < j_hi; ++j) {
            const int dst = s_dst[j];
            const unsigned src = s_src[j];
            int rx = -1, ry = -1, rp = -1;
            if (p.mode != 0) { rx = s_row[3 * j]; ry = s_row[3 * j + 1]; rp = s_row[3 * j + 2]; }
            const double* s = acc + (size_t)(src & 255u) * 128 + lane;
            double v0 = s[0], v1 = s[32], v2 = s[64], v3 = s[96];
#pragma unroll
            for (int k = 1; k < 4; ++k) {
                const unsigned rk = (src >> (8 * k)) & 255u;
                if (rk != 255u) {  // warp-uniform
                    const double* t = acc + (size_t)rk * 128 + lane;
                    v0 += t[0]; v1 += t[32]; v2 += t[64]; v3 += t[96];
                }
            }
            if (dst >= 0) {
                emit_node<false>(p.mode, p.bprev, p.Zb, rx, ry, rp, p.a, p.b, p.nN, ldb, dst, b, v0, v1, v2, v3);
            } else {
                double* z = p.scratch + (size_t)(-1 - dst) * 4 * ldb + b;
                z[0] = v0; z[ldb] = v1; z[2 * ldb] = v2; z[3 * ldb] = v3;
            }
        }
    }
    __shared__ double se[EP_WARPS][32];
    se[w][lane] = e_acc;
    __syncthreads();
    if (w == 0) {
        double t = se[0][lane];
#pragma unroll
        for (int k = 1; k < EP_WARPS; ++k) t += se[k][lane];
        p.epart[(size_t)blockIdx.x * ldb + b] = t;
    }
}

// nodes shared between patches: a/b rows = sum of the patches' partials (fixed order).
// grid = (ceil(nshared/8), ldb/32), block = (32, 8)
__global__ void __launch_bounds__(256) k_patch_merge(int nshared, const int* __restrict__ mptr, const int* __restrict__ msrc,
                                                    const int* __restrict__ mnode, const int* __restrict__ mrow,
                                                    const double* __restrict__ scratch, double* a, double* bvec,
                                                    const double* bprev, double* Zb, int mode, int nN, int ldb) {
    const int i = blockIdx.x * blockDim.y + threadIdx.y;
    const int b = blockIdx.y * 32 + threadIdx.x;
    if (i >= nshared) return;
    const size_t L = (size_t)ldb;
    const int k0 = __ldg(mptr + i), k1 = __ldg(mptr + i + 1), nd = __ldg(mnode + i);
    const int rx = __ldg(mrow + 3 * i), ry = __ldg(mrow + 3 * i + 1), rp = __ldg(mrow + 3 * i + 2);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    for (int k = k0; k < k1; ++k) {
        const double* z = scratch + (size_t)__ldg(msrc + k) * 4 * L + b;
        s0 += z[0]; s1 += z[L]; s2 += z[2 * L]; s3 += z[3 * L];
    }
    emit_node<true>(mode, bprev, Zb, rx, ry, rp, a, bvec, nN, L, nd, b, s0, s1, s2, s3);
}

// CSR SpMM over the ensemble's multi-RHS block: Z[r] = sum_j val[j] * X[idx[j]] for the n solver rows (Crank-Nicolson explicit
// operator E = M/dt - (C + D + K/Re)/2 applied to u_n).  One warp = one row x 32 trajectories: column indices and values
// are warp-uniform (broadcast), every X access is a coalesced 256-byte row segment, consecutive rows are neighbours in
// the mesh (nested-dissection order) and share most of their columns through L1.  grid = (ceil(n/8), ldb/32), block = (32, 8)
__global__ void __launch_bounds__(256) k_spmm(int n, const int* __restrict__ ptr, const int* __restrict__ idx,
                                             const double* __restrict__ val, const double* __restrict__ X, double* __restrict__ Z,
                                             int ldb) {
    const int r = blockIdx.x * blockDim.y + threadIdx.y;
    const int b = blockIdx.y * 32 + threadIdx.x;
    if (r >= n) return;
    const int j0 = __ldg(ptr + r), j1 = __ldg(ptr + r + 1);
    double s0 = 0.0, s1 = 0.0;
    int j = j0;
    for (; j + 1 < j1; j += 2) {
        s0 = fma(__ldg(val + j), X[(size_t)__ldg(idx + j) * ldb + b], s0);
        s1 = fma(__ldg(val + j + 1), X[(size_t)__ldg(idx + j + 1) * ldb + b], s1);
    }
    if (j < j1) s0 = fma(__ldg(val + j), X[(size_t)__ldg(idx + j) * ldb + b], s0);
    Z[(size_t)r * ldb + b] = s0 + s1;
}

// Z[dst[i]] = sum of Z[src[j]] (fixed order): right-hand side of the merged top of the elimination tree.
// grid = (ceil(n/8), ldb/32), block = (32, 8)
__global__ void __launch_bounds__(256) k_gather_sum(int nrows, const int* __restrict__ ptr, const int* __restrict__ src,
                                                   const int* __restrict__ dst, double* Z, int ldb) {
    const int i = blockIdx.x * blockDim.y + threadIdx.y;
    const int b = blockIdx.y * 32 + threadIdx.x;
    // launched between two sweep launches with programmatic stream serialisation: the next launch may start its prologue
    // now; nothing of Z is touched before the previous launch has completed (every thread waits, so that the completion
    // of this grid implies the completion of the one before it)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    int k0 = 0, k1 = 0, d = 0;
    if (i < nrows) { k0 = __ldg(ptr + i); k1 = __ldg(ptr + i + 1); d = __ldg(dst + i); }
    int s0 = -1, s1 = -1, s2 = -1;  // the first three sources (a y row: b and two update vectors) are fetched together
    if (k0 < k1) s0 = __ldg(src + k0);
    if (k0 + 1 < k1) s1 = __ldg(src + k0 + 1);
    if (k0 + 2 < k1) s2 = __ldg(src + k0 + 2);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (i >= nrows) return;
    const double v0 = s0 >= 0 ? Z[(size_t)s0 * ldb + b] : 0.0;
    const double v1 = s1 >= 0 ? Z[(size_t)s1 * ldb + b] : 0.0;
    const double v2 = s2 >= 0 ? Z[(size_t)s2 * ldb + b] : 0.0;
    double s = 0.0;
    if (s0 >= 0) s += v0;
    if (s1 >= 0) s += v1;
    if (s2 >= 0) s += v2;
    for (int k = k0 + 3; k < k1; ++k) s += Z[(size_t)__ldg(src + k) * ldb + b];
    Z[(size_t)d * ldb + b] = s;
}

// control terms of the fused right-hand side: Z[row] += sum_k coef[k][i] * u_ctrl[k] on the (few) rows where the lifting /
// force vectors are non-zero.  grid = (ceil(nrows/8), ldb/32), block = (32, 8)
__global__ void __launch_bounds__(256) k_ctrl_add(int nrows, const int* __restrict__ rows, const double* __restrict__ coef, int na,
                                                 const double* __restrict__ uctrl, double* __restrict__ Z, int ldb) {
    const int i = blockIdx.x * blockDim.y + threadIdx.y;
    const int b = blockIdx.y * 32 + threadIdx.x;
    if (i >= nrows) return;
    double v = Z[(size_t)__ldg(rows + i) * ldb + b];
    for (int k = 0; k < na; ++k) v = fma(__ldg(coef + (size_t)k * nrows + i), uctrl[(size_t)k * ldb + b], v);
    Z[(size_t)__ldg(rows + i) * ldb + b] = v;
}

// Dirichlet rows of the new state: value = sum_k shape[k][j] * u_ctrl[k].  grid = (ceil(nbc/8), ldb/32), block = (32, 8)
__global__ void __launch_bounds__(256) k_bc_fill(int nbc, const int* __restrict__ bc_dofs, int na, const double* __restrict__ bc_shape,
                                                const double* __restrict__ uctrl, double* __restrict__ up, int ldb) {
    const int j = blockIdx.x * blockDim.y + threadIdx.y;
    const int b = blockIdx.y * 32 + threadIdx.x;
    if (j >= nbc) return;
    double v = 0.0;
    for (int k = 0; k < na; ++k) {
        const double sh = __ldg(bc_shape + (size_t)k * nbc + j);
        if (sh != 0.0) v = fma(sh, uctrl[(size_t)k * ldb + b], v);
    }
    up[(size_t)__ldg(bc_dofs + j) * ldb + b] = v;
}

// rhs in solver row order.  grid = (ceil(n/8), ldb/32), block = (32, 8)
__global__ void __launch_bounds__(256) k_rhs_build(int n, int Nv, const int* __restrict__ perm,
                                                  const double* __restrict__ a, const double* __restrict__ bprev,
                                                  int order, int na, const double* __restrict__ ctrl_rhs,
                                                  const double* __restrict__ uctrl, double* __restrict__ Z, int ldb) {
    const int r = blockIdx.x * blockDim.y + threadIdx.y;
    const int b = blockIdx.y * 32 + threadIdx.x;
    if (r >= n) return;
    const int dof = __ldg(perm + r);
    double v = 0.0;
    if (dof < Nv) {
        const size_t o = (size_t)dof * ldb + b;
        v = (order == 2) ? (a[o] + bprev[o]) : 0.5 * a[o];
    }
    for (int k = 0; k < na; ++k) {
        const double c = __ldg(ctrl_rhs + (size_t)k * n + r);
        if (c != 0.0) v = fma(c, uctrl[(size_t)k * ldb + b], v);
    }
    Z[(size_t)r * ldb + b] = v;
}

// ----------------------------------------------------------------------------------------------
// Multifrontal sweeps: dense tile products on the FP64 tensor cores, fed by bulk-async (TMA) copies.
//
// One job (a row tile of a SolvePlan block) = 8*nrb output rows x the W = 32*NWC trajectories of the CTA's slab:
//   x_k   = Z[i0[k]] (+ Z[i1[k]] + Z[i2[k]])          gathered input rows
//   acc   = sum_k V[:,k] x_k                           mma.sync m8n8k4 f64: M = rows, N = trajectories
//   Z[out0 + r] = acc_r (+ Z[e0[r]] + Z[e1[r]])        r < nr
// CTA = NWC consumer warps (32 trajectories each) + 1 producer warp.  Per stage of a ring the
// producer copies up to SV_SLOTS gathered rows and the matching slice of V into shared memory and
// signals an mbarrier; consumers multiply out of shared memory and hand the stage back through a
// second mbarrier.  The gather lists of a multifrontal solve are almost entirely runs of consecutive
// rows, and Z is described to the TMA unit as a 2-D tensor [rows][ldb]: a run of 1/2/4/8 rows x W
// columns is ONE cp.async.bulk.tensor box copy (1 to 4 copies per stage instead of 12 row copies), it
// lands densely in shared memory, and rows past the end of Z read as zeros (absent gathers).
// V is packed on the host in A-fragment order (a fragment = one conflict-free 256-byte read).
// B fragments are read as 16-byte pairs: lane (g = lane/4, t = lane%4) reads columns 2g, 2g+1 of row
// k0+t in a 16-column group and feeds two MMAs (n-block "even columns", n-block "odd columns").  The C
// fragments of such a pair hold 4 consecutive trajectories per lane (32-byte stores).
// CTAs are persistent: the host compiles each launch into one instruction stream per CTA, and the
// ring keeps filling across job boundaries.  Every Z row is written by exactly one job: no atomics.
// ----------------------------------------------------------------------------------------------
#ifndef SV_MINCTAS4
#define SV_MINCTAS4 3  // CTAs per SM the 4-warp sweep kernels are compiled for (3 => 128 registers per thread)
#endif
#ifndef SV_MINCTAS8
#define SV_MINCTAS8 2  // CTAs per SM the 8-warp sweep kernel is compiled for (2 => 96 registers per thread)
#endif
constexpr int SV_MAXSTAGES = 8;  // ring depth is chosen per launch from the shared memory one CTA may use
constexpr int SV_PREFETCH = 4;   // stage records the producer keeps in flight
constexpr int SV_JREC = 72;      // ints per job record
constexpr int SV_SREC = 32;      // ints per stage record
constexpr int SV_MAXRUNS = 13;   // box copies per stage record
constexpr int SV_RED_BYTES = 3 * 32 * 32 * 8;  // k-split: partial accumulators of consumer warps 1..3

// Shared-memory geometry of one ring stage for NWC consumer warps (W = 32*NWC trajectories) and `slots`
// gathered rows per stage (a launch parameter: 3 planes x slots/3 k for nsrc == 3, slots k otherwise).
template <int NWC>
struct SweepCfg {
    static constexpr int W = 32 * NWC;  // trajectories per CTA
    static constexpr int XS = W + 4;    // shared-memory row stride in doubles (== 4 mod 16: conflict-free B fragments)
    static constexpr int MIN_CTAS = NWC == 8 ? SV_MINCTAS8 : (NWC == 4 ? SV_MINCTAS4 : 4);
    __host__ __device__ static constexpr int voff(int slots) { return slots * XS * 8; }            // V slice: slots k x 32 rows
    __host__ __device__ static constexpr int joff(int slots) { return voff(slots) + slots * 256; }  // job record (first stage of a job)
    __host__ __device__ static constexpr int hoff(int slots) { return joff(slots) + 384; }          // stage header (nk, ...)
    __host__ __device__ static constexpr int stage_bytes(int slots) { return hoff(slots) + 128; }
    __host__ __device__ static constexpr int smem_bytes(int stages, int slots) { return stages * stage_bytes(slots) + 2 * SV_MAXSTAGES * 8; }
};

// The host compiles every launch into one instruction stream per CTA (upload_plan):
//   stage record (32 ints): nk, npieces, V offset / 32 doubles, V bytes, job record index if this is the first
//                           stage of a job (else -1), total rows, then per piece (source row, slot << 8 | log2 rows);
//                           its first 16 bytes also ride into shared memory as the stage header (consumers read nk)
//   job record   (72 ints): K, nrb, nr, nsrc, out0, ystore, seed mode, stages, e0[32], e1[32]
//                           (seed mode 1: e0/e1 = update rows to start from; 2: e0 = canonical dof of each output row;
//                           3: as 2, and the rows are NOT stored in solver order)
// The producer reads consecutive 128-byte stage records; the job record rides into shared memory with
// the job's first stage, so nothing on the device chases a pointer.

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// one box of the Z tensor: rows [row, row + box height) x columns [col, col + box width)
__device__ __forceinline__ void tma_box_g2s(uint32_t dst, const void* tmap, int col, int row, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(tmap), "r"(col), "r"(row), "r"(bar)
        : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}


// ---- block-sparse SpMM on the FP64 tensor cores (Crank-Nicolson explicit operator E u_n) --------------------------
// Z[row] = sum_j E[row, j] X[j] for the free velocity rows, all trajectories.  The library re-packs the caller's CSR
// (build_spmm): consecutive solver rows (a compact piece of the mesh in the nested-dissection order) form a block of
// at most 64 rows whose distinct columns (<= SPMM_CMAX) are staged once in shared memory, 32 trajectories wide; inside a
// block, rows with similar column sets are grouped by eight, and each group is stored as a dense 8 x K panel over the
// union of its columns, in mma.m8n8k4 A-fragment order.  One CTA = one (block, 32-trajectory slice), one warp = one
// group: per k-step one A fragment (a coalesced 256-byte load, prefetched four k-steps ahead in registers), four column
// slots, four B fragments from shared memory and four DMMAs (8 rows x 4 columns x 32 trajectories).  All index loads
// are addressed by the block number alone (fixed strides), so the two dependent chains (columns -> X rows, group
// record -> panel) start together; four CTAs per SM overlap one another's gathers and DMMAs.
// History (cylinder, 256 trajectories): scalar CSR row loop (k_spmm below, kept for A/B runs) 0.21 ms - one 256-byte
// row segment of X per non-zero, bound by L1 wavefronts; persistent CTAs with a two-stage mbarrier ring 0.28 ms with
// per-row bulk copies (issued lane by lane through the uniform datapath), 0.135 ms with a cp.async producer warp,
// 0.115 ms with every warp gathering - one CTA per SM leaves too little slack around the per-item barrier.
#ifndef FCB_SPMM_CMAX
#define FCB_SPMM_CMAX 192
#endif
#ifndef FCB_SPMM_MINCTAS
#define FCB_SPMM_MINCTAS 4
#endif
constexpr int SPMM_CMAX = FCB_SPMM_CMAX;   // staged columns per block (a multiple of 64)
constexpr int SPMM_XS = 36;      // shared row stride in doubles (== 4 mod 16: k-steps with slots distinct mod 4 are conflict-free)
constexpr int SPMM_WARPS = 8;    // warps = groups per block
constexpr int SPMM_GREC = 12;    // ints per group record: first k-step, k-steps, 8 solver rows (-1 = none), 2 pad
constexpr int SPMM_SMEM = SPMM_CMAX * SPMM_XS * 8;

struct SpmmArgs {
    const int* ucols;              // [nblk][SPMM_CMAX] staged column (canonical velocity dof) per slot, -1 = none
    const int* ginfo;              // [nblk][SPMM_WARPS][SPMM_GREC]
    const unsigned short* kslots;  // [nk/4][4 columns][4 k-steps] slot of each k-step's columns, chunked by four k-steps
    const double* avals;           // [nk][32] panel values in A-fragment order: lane l holds (row l/4, column l%4)
    const double* X;
    double* Z;
    int ldb, nslice;
};

__global__ void __launch_bounds__(32 * SPMM_WARPS, FCB_SPMM_MINCTAS) k_spmm_mma(const SpmmArgs p) {
    extern __shared__ __align__(16) double xs[];  // [slot][SPMM_XS]
    const int lane = threadIdx.x, w = threadIdx.y;
    const int blk = blockIdx.x / p.nslice, b0 = (blockIdx.x - blk * p.nslice) * 32;
    const size_t ldb = (size_t)p.ldb;
    const int half = lane >> 4, l16 = lane & 15, kq = lane & 3, nq = lane >> 2;
    // this half-warp gathers the staged columns c = 2 (w + 8 i) + half, i = 0..11; lane l16 = i holds the index of round i
    const int c_mine = 2 * (w + SPMM_WARPS * l16) + half;
    const int col = c_mine < SPMM_CMAX ? __ldg(p.ucols + (size_t)blk * SPMM_CMAX + c_mine) : -1;
    const int* gi = p.ginfo + ((size_t)blk * SPMM_WARPS + w) * SPMM_GREC;
    const int k0 = __ldg(gi), nk = __ldg(gi + 1), row = __ldg(gi + 2 + nq);
#pragma unroll
    for (int i = 0; i < SPMM_CMAX / (2 * SPMM_WARPS); ++i) {
        const int cc = __shfl_sync(0xffffffffu, col, (lane & 16) + i);
        if (cc >= 0) {
            const uint32_t dst = smem_u32(xs + (size_t)(2 * (w + SPMM_WARPS * i) + half) * SPMM_XS + 2 * l16);
            const double* src = p.X + (size_t)cc * ldb + b0 + 2 * l16;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    // panel prefetch: chunks of four k-steps (four coalesced 256-byte A loads + one 8-byte load of the lane's four
    // slots), three chunks in registers - two are in flight while one is multiplied.  Groups start on chunk boundaries
    // and the arrays are padded, so the unconditional loads of the first two chunks stay inside.
    const double* av = p.avals + (size_t)k0 * 32 + lane;
    const uint2* sl = reinterpret_cast<const uint2*>(p.kslots) + (size_t)k0 + kq;  // [chunk][kq][4] u16
    const int nch = (nk + 3) >> 2;
    double a_buf[3][4];
    uint2 s_buf[3];
#define SPMM_LOAD(B, C)                                                         \
    do {                                                                        \
        _Pragma("unroll") for (int t = 0; t < 4; ++t) a_buf[B][t] = __ldg(av + ((C) * 4 + t) * 32); \
        s_buf[B] = __ldg(sl + (C) * 4);                                         \
    } while (0)
    SPMM_LOAD(0, 0);
    SPMM_LOAD(1, 1);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    double acc[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[j][0] = 0.0; acc[j][1] = 0.0; }
#define SPMM_CHUNK(B, C)                                                        \
    do {                                                                        \
        _Pragma("unroll") for (int t = 0; t < 4; ++t) {                         \
            if ((C) * 4 + t < nk) {                                             \
                const unsigned int pair = t < 2 ? s_buf[B].x : s_buf[B].y;      \
                const double* xr = xs + ((pair >> (16 * (t & 1))) & 0xffffu) * SPMM_XS + nq; \
                _Pragma("unroll") for (int j = 0; j < 4; ++j) dmma(acc[j], a_buf[B][t], xr[8 * j]); \
            }                                                                   \
        }                                                                       \
    } while (0)
    for (int c = 0; c < nch; c += 3) {
        if (c + 2 < nch) SPMM_LOAD(2, c + 2);
        SPMM_CHUNK(0, c);
        if (c + 3 < nch) SPMM_LOAD(0, c + 3);
        if (c + 1 < nch) SPMM_CHUNK(1, c + 1);
        if (c + 4 < nch) SPMM_LOAD(1, c + 4);
        if (c + 2 < nch) SPMM_CHUNK(2, c + 2);
    }
#undef SPMM_LOAD
#undef SPMM_CHUNK
    if (row >= 0) {
        double* z = p.Z + (size_t)row * ldb + b0 + 2 * kq;
#pragma unroll
        for (int j = 0; j < 4; ++j) *reinterpret_cast<double2*>(z + 8 * j) = make_double2(acc[j][0], acc[j][1]);
    }
}

struct alignas(64) SweepMaps {
    CUtensorMap m[6];  // boxes of 1, 2, 4, 8, 16, 32 rows x W columns of Z
};

struct RingPos {
    int s;
    uint32_t ph;
    __device__ __forceinline__ void advance(int nstages) {
        if (++s == nstages) { s = 0; ph ^= 1; }
    }
};

// One job on one consumer warp: seeds, the stage loop (NRB row blocks x 2 pairs of trajectory blocks), store.
// acc[rb][p][0..3] = trajectories t0 + 16p + 4*(lane%4) + 0..3 of output row 8*rb + lane/4.
// KS (k-split, narrow CTAs only): the 4 consumer warps share ONE 32-trajectory tile and take the job's stages
// round-robin (warp kw owns stages kw, kw+4, ...); partial sums meet in shared memory and warp 0 stores.
#ifndef FCB_KS_REDUCER
#define FCB_KS_REDUCER 3
#endif
#ifndef FCB_SWEEP_SLIM
#define FCB_SWEEP_SLIM 0  // 1: the three-plane gather and the y store become run-time branches (5 job variants per kernel instead of 20:
                          // a third of the code; measured on B200: forward 2 % faster, backward 4 % slower, step +0.6 %)
#endif
#if FCB_SWEEP_SLIM
template <int NWC, bool KS, int NRB>
__device__ __forceinline__ void sweep_job(unsigned char* smem, uint32_t bar_full, uint32_t bar_empty, RingPos& rp, int nstages,
                                          int slots, const int* jh, int nst, int K, int nr, int out0, int ystore, int seed,
                                          double* Z, size_t L, int t0, int c0, int lane, int kw, double* red, double* xout,
                                          int* diverged, int Nv, const bool SRC3, const bool YST) {
#else
template <int NWC, bool KS, int NRB, bool SRC3, bool YST>
__device__ __forceinline__ void sweep_job(unsigned char* smem, uint32_t bar_full, uint32_t bar_empty, RingPos& rp, int nstages,
                                          int slots, const int* jh, int nst, int K, int nr, int out0, int ystore, int seed,
                                          double* Z, size_t L, int t0, int c0, int lane, int kw, double* red, double* xout,
                                          int* diverged, int Nv) {
#endif
    // t0: first trajectory (global column) of this warp; c0: its first column inside the CTA's shared-memory rows
    using C = SweepCfg<NWC>;
    constexpr int XS = C::XS;
    constexpr int NA = NRB > 0 ? NRB : 1;
    const int gid = lane >> 2, tig = lane & 3;
    const int stage_bytes = C::stage_bytes(slots), voff = C::voff(slots), hoff = C::hoff(slots);
    const int plane = SRC3 ? ((slots / 3) & ~3) * XS : 0;  // doubles between the three source planes of a stage
    double ce[NA][2][2], co[NA][2][2];  // even / odd column MMAs of each pair
#pragma unroll
    for (int rb = 0; rb < NA; ++rb)
#pragma unroll
        for (int p = 0; p < 2; ++p) ce[rb][p][0] = ce[rb][p][1] = co[rb][p][0] = co[rb][p][1] = 0.0;
    int xdof[NA];  // seed mode 2: canonical dof of this lane's output rows (read now: the record's stage slot is recycled later)
#pragma unroll
    for (int rb = 0; rb < NA; ++rb) xdof[rb] = (NRB > 0 && seed >= 2) ? jh[8 + rb * 8 + gid] : 0;
    // k-split: warp 3 owns the fewest stages (3, 7, ...: none at all in a job of up to three stages), so it is the one that
    // waits for the seed rows, collects the partial sums and stores (FCB_KS_REDUCER picks the warp; 0 = the first)
    constexpr int RW = KS ? FCB_KS_REDUCER : 0;
    if (NRB > 0 && seed == 1 && (!KS || kw == RW)) {
        // the children's update rows that land on these output rows seed the accumulators
#pragma unroll
        for (int rb = 0; rb < NRB; ++rb) {
            const int ea = jh[8 + rb * 8 + gid], eb = jh[40 + rb * 8 + gid];
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                const int col = t0 + 16 * p + 4 * tig;
                if (ea >= 0) {
                    const double4 v = *reinterpret_cast<const double4*>(Z + (size_t)ea * L + col);
                    ce[rb][p][0] += v.x; co[rb][p][0] += v.y; ce[rb][p][1] += v.z; co[rb][p][1] += v.w;
                }
                if (eb >= 0) {
                    const double4 v = *reinterpret_cast<const double4*>(Z + (size_t)eb * L + col);
                    ce[rb][p][0] += v.x; co[rb][p][0] += v.y; ce[rb][p][1] += v.z; co[rb][p][1] += v.w;
                }
            }
        }
    }
    double* zy = YST ? Z + (size_t)(ystore + tig) * L + t0 + 2 * gid : nullptr;
    int kleft = K - tig;  // rows of this lane's k index still inside the job
    for (int si = 0; si < nst; ++si) {
        mbar_wait(bar_full + 8 * rp.s, rp.ph);
        const unsigned char* st = smem + rp.s * stage_bytes;
        const int nk = *reinterpret_cast<const int*>(st + hoff);
        const double* xr = reinterpret_cast<const double*>(st) + tig * XS + c0 + 2 * gid;
        const double* vs = reinterpret_cast<const double*>(st + voff) + lane;
        if (KS && (si & 3) != kw) {  // another warp's stage: only keep pace with the ring
            if (YST) { zy += (size_t)nk * L; kleft -= nk; }
        } else
        for (int g = 0; g < nk; g += 4) {
            double a[NA];
#pragma unroll
            for (int rb = 0; rb < NRB; ++rb) a[rb] = vs[rb * 32];
#pragma unroll
            for (int p = 0; p < 2; ++p) {
                double2 bv = *reinterpret_cast<const double2*>(xr + 16 * p);
                if (SRC3) {
                    const double2 b1 = *reinterpret_cast<const double2*>(xr + plane + 16 * p);
                    const double2 b2 = *reinterpret_cast<const double2*>(xr + 2 * plane + 16 * p);
                    bv.x += b1.x + b2.x;
                    bv.y += b1.y + b2.y;
                }
                if (YST && kleft > 0) *reinterpret_cast<double2*>(zy + 16 * p) = bv;
#pragma unroll
                for (int rb = 0; rb < NRB; ++rb) {
                    dmma(ce[rb][p], a[rb], bv.x);
                    dmma(co[rb][p], a[rb], bv.y);
                }
            }
            vs += NRB * 32;
            xr += 4 * XS;
            if (YST) { zy += 4 * L; kleft -= 4; }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty + 8 * rp.s);
        rp.advance(nstages);
    }
    if (KS && NRB > 0) {
        // k-split: warps 1..3 park their partial sums in shared memory, warp 0 adds them up
        asm volatile("bar.sync 1, 128;" ::: "memory");  // warp 0 is done reading the previous job's partials
        if (kw != RW) {
            double* r = red + (size_t)(kw < RW ? kw : kw - 1) * (NRB * 8 * 32) + lane;
#pragma unroll
            for (int rb = 0; rb < NRB; ++rb)
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    r[((rb * 2 + p) * 4 + 0) * 32] = ce[rb][p][0];
                    r[((rb * 2 + p) * 4 + 1) * 32] = ce[rb][p][1];
                    r[((rb * 2 + p) * 4 + 2) * 32] = co[rb][p][0];
                    r[((rb * 2 + p) * 4 + 3) * 32] = co[rb][p][1];
                }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (kw != RW) return;
#pragma unroll
        for (int w = 0; w < 3; ++w) {
            const double* r = red + (size_t)w * (NRB * 8 * 32) + lane;
#pragma unroll
            for (int rb = 0; rb < NRB; ++rb)
#pragma unroll
                for (int p = 0; p < 2; ++p) {
                    ce[rb][p][0] += r[((rb * 2 + p) * 4 + 0) * 32];
                    ce[rb][p][1] += r[((rb * 2 + p) * 4 + 1) * 32];
                    co[rb][p][0] += r[((rb * 2 + p) * 4 + 2) * 32];
                    co[rb][p][1] += r[((rb * 2 + p) * 4 + 3) * 32];
                }
        }
    }
    if (NRB > 0) {
#pragma unroll
        for (int rb = 0; rb < NRB; ++rb) {
            const int r = rb * 8 + gid;
            if (r < nr) {
                double* zo = Z + (size_t)(out0 + r) * L + t0 + 4 * tig;
                if (seed != 3) {  // mode 3: solution rows that no other block gathers are not kept in solver order
#pragma unroll
                    for (int p = 0; p < 2; ++p)
                        *reinterpret_cast<double4*>(zo + 16 * p) = make_double4(ce[rb][p][0], co[rb][p][0], ce[rb][p][1], co[rb][p][1]);
                }
                if (seed >= 2) {
                    // backward sweep: x_t also goes out in canonical numbering (the new state), with the divergence check
                    const int dof = xdof[rb];
                    double* xo = xout + (size_t)dof * L + t0 + 4 * tig;
#pragma unroll
                    for (int p = 0; p < 2; ++p) {
                        const double4 v = make_double4(ce[rb][p][0], co[rb][p][0], ce[rb][p][1], co[rb][p][1]);
                        *reinterpret_cast<double4*>(xo + 16 * p) = v;
                        if (dof < Nv) {
                            int* dv = diverged + t0 + 16 * p + 4 * tig;
                            if (!isfinite(v.x)) dv[0] = 1;
                            if (!isfinite(v.y)) dv[1] = 1;
                            if (!isfinite(v.z)) dv[2] = 1;
                            if (!isfinite(v.w)) dv[3] = 1;
                        }
                    }
                }
            }
        }
    }
}

// grid = (persistent CTAs, slabs of 32*NWC trajectories); ldb is a multiple of 32*NWC.
// block = (32, NWC + 1), or (32, 5) for the k-split variant (NWC == 1, four consumer warps on one tile)
template <int NWC, bool KS>
__global__ void __launch_bounds__(32 * (KS ? 5 : NWC + 1), KS ? SV_MINCTAS4 : SweepCfg<NWC>::MIN_CTAS)
    k_front_sweep(const __grid_constant__ SweepMaps maps, const int* __restrict__ srec, const int* __restrict__ jrec,
                  const int* __restrict__ cta_sptr, const int* __restrict__ cta_jptr, const double* __restrict__ vals,
                  double* Z, int ldb, int nstages, int slots, double* xout, int* diverged, int Nv, unsigned long long* dbg) {
    using C = SweepCfg<NWC>;
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int NCW = KS ? 4 : NWC;  // consumer warps
    const int lane = threadIdx.x, wid = threadIdx.y;
    const int slab0 = blockIdx.y * C::W;
    if (dbg) {  // FCB_SWEEP_DEBUG timeline: 8 slots per CTA (see fcb_profile_step)
        dbg += (size_t)(blockIdx.y * gridDim.x + blockIdx.x) * 8;
        if (lane == 0 && wid == 0) dbg[0] = globaltimer_ns();
    }
    const uint32_t sbase = smem_u32(smem);
    const int stage_bytes = C::stage_bytes(slots);
    const uint32_t bar_full = sbase + nstages * stage_bytes;
    const uint32_t bar_empty = bar_full + SV_MAXSTAGES * 8;
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        for (int s = 0; s < nstages; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, NCW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // the next level may start its prologue

    if (wid == NCW) {
        // ---------------- producer warp: stream the stage records, issue the bulk copies ----------------
        const int sbeg = __ldg(cta_sptr + blockIdx.x), ns = __ldg(cta_sptr + blockIdx.x + 1) - sbeg;
        const int* rec = srec + (size_t)sbeg * SV_SREC + lane;
        int q[SV_PREFETCH];
#pragma unroll
        for (int u = 0; u < SV_PREFETCH; ++u) q[u] = (u < ns) ? __ldg(rec + u * SV_SREC) : 0;
        RingPos rp{0, 1};  // parity 1: the first wait on a fresh "empty" barrier passes
        asm volatile("griddepcontrol.wait;" ::: "memory");  // Z rows of the previous level are complete and visible
        for (int c0 = 0; c0 < ns; c0 += SV_PREFETCH) {
            int qn[SV_PREFETCH];
#pragma unroll
            for (int u = 0; u < SV_PREFETCH; ++u)
                qn[u] = (c0 + SV_PREFETCH + u < ns) ? __ldg(rec + (c0 + SV_PREFETCH + u) * SV_SREC) : 0;
#pragma unroll
            for (int u = 0; u < SV_PREFETCH; ++u) {
                if (c0 + u < ns) {  // warp-uniform
                    const int v = q[u];
                    const int npieces = __shfl_sync(0xffffffffu, v, 1);
                    const int voff = __shfl_sync(0xffffffffu, v, 2);
                    const uint32_t vbytes = (uint32_t)__shfl_sync(0xffffffffu, v, 3);
                    const int jidx = __shfl_sync(0xffffffffu, v, 4);
                    const uint32_t xbytes = (uint32_t)__shfl_sync(0xffffffffu, v, 5);
                    const int src = __shfl_sync(0xffffffffu, v, 6 + 2 * (lane < SV_MAXRUNS ? lane : 0));
                    const int dsc = __shfl_sync(0xffffffffu, v, 7 + 2 * (lane < SV_MAXRUNS ? lane : 0));
                    const uint32_t full = bar_full + 8 * rp.s;
                    const uint32_t sx = sbase + rp.s * stage_bytes;
                    mbar_wait(bar_empty + 8 * rp.s, rp.ph);
                    if (lane == 0) {
                        mbar_expect_tx(full, xbytes + vbytes + 16u + (jidx >= 0 ? SV_JREC * 4u : 0u));
                        bulk_g2s(sx + C::hoff(slots), srec + (size_t)(sbeg + c0 + u) * SV_SREC, 16u, full);
                        if (vbytes) bulk_g2s(sx + C::voff(slots), vals + (size_t)voff * 32, vbytes, full);
                        if (jidx >= 0) bulk_g2s(sx + C::joff(slots), jrec + (size_t)jidx * SV_JREC, SV_JREC * 4u, full);
                    }
                    __syncwarp();
                    if (lane < npieces) {
                        const uint32_t dst = sx + (uint32_t)((dsc >> 8) * C::XS * 8);
                        if (dsc & 128)  // a single row (its 4-row group is not a run of consecutive rows)
                            bulk_g2s(dst, Z + (size_t)src * ldb + slab0, (uint32_t)(C::W * 8), full);
                        else  // a run of 4..32 consecutive rows, starting on a 4-row boundary: one box copy
                            tma_box_g2s(dst, &maps.m[dsc & 7], slab0, src, full);
                    }
                    if (dbg && lane == 0 && c0 + u == 0) dbg[1] = globaltimer_ns();
                    rp.advance(nstages);
                }
            }
#pragma unroll
            for (int u = 0; u < SV_PREFETCH; ++u) q[u] = qn[u];
        }
        if (dbg && lane == 0) { dbg[5] = globaltimer_ns(); dbg[7] = (unsigned long long)ns; }
        return;
    }

    // ---------------- consumer warps ----------------
    const int c0 = KS ? 0 : wid * 32, t0 = slab0 + c0;  // first column / trajectory of this warp
    double* red = reinterpret_cast<double*>(smem + nstages * stage_bytes + 2 * SV_MAXSTAGES * 8);  // k-split partial sums
    const size_t L = (size_t)ldb;
    const int nj = __ldg(cta_jptr + blockIdx.x + 1) - __ldg(cta_jptr + blockIdx.x);
    RingPos rp{0, 0};
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int j = 0; j < nj; ++j) {
        mbar_wait(bar_full + 8 * rp.s, rp.ph);  // the job record arrives with the job's first stage
        if (dbg && j == 0 && lane == 0 && wid == 0) dbg[2] = globaltimer_ns();
        const int* jh = reinterpret_cast<const int*>(smem + rp.s * stage_bytes + C::joff(slots));
        const int4 h0 = *reinterpret_cast<const int4*>(jh);
        const int4 h1 = *reinterpret_cast<const int4*>(jh + 4);
        const int K = h0.x, nrb = h0.y, nr = h0.z, out0 = h1.x, ystore = h1.y, nst = h1.w;
        const bool src3 = h0.w == 3;
        const int seed = h1.z;  // 0: none, 1: accumulators start from the children's update rows, 2: outputs also go to xout
#if FCB_SWEEP_SLIM
#define SWEEP_CASE(NRB)                                                                                                              \
    case NRB:                                                                                                                        \
        sweep_job<NWC, KS, NRB>(smem, bar_full, bar_empty, rp, nstages, slots, jh, nst, K, nr, out0, ystore, seed, Z, L, t0, c0, lane, \
                                wid, red, xout, diverged, Nv, src3, ystore >= 0);                                                    \
        break;
#else
#define SWEEP_RUN(NRB, S3, YS) \
    sweep_job<NWC, KS, NRB, S3, YS>(smem, bar_full, bar_empty, rp, nstages, slots, jh, nst, K, nr, out0, ystore, seed, Z, L, t0, c0, lane, wid, red, xout, diverged, Nv)
#define SWEEP_CASE(NRB)                                        \
    case NRB:                                                  \
        if (src3) {                                            \
            if (ystore >= 0) SWEEP_RUN(NRB, true, true);       \
            else SWEEP_RUN(NRB, true, false);                  \
        } else {                                               \
            if (ystore >= 0) SWEEP_RUN(NRB, false, true);      \
            else SWEEP_RUN(NRB, false, false);                 \
        }                                                      \
        break;
#endif
        switch (nrb) {
            SWEEP_CASE(0)
            SWEEP_CASE(1)
            SWEEP_CASE(2)
            SWEEP_CASE(3)
            SWEEP_CASE(4)
        }
#undef SWEEP_CASE
#if !FCB_SWEEP_SLIM
#undef SWEEP_RUN
#endif
        if (dbg && j == 0 && lane == 0 && wid == 0) dbg[3] = globaltimer_ns();
    }
    if (dbg && lane == 0 && wid == 0) { dbg[4] = globaltimer_ns(); dbg[6] = (unsigned long long)nj; }
}


// ----------------------------------------------------------------------------------------------
// Subtree clusters: the lower part of the elimination tree, swept with everything resident in shared memory.
//
// A cluster (multifrontal.py: choose_clusters) is a connected piece of the tree; one CTA sweeps it for 32 trajectories on
// one resident vector S = [own unknowns of its fronts | boundary rows of its root] (row stride 36 doubles, so that the
// B fragments below are bank-conflict free for consecutive rows):
//   forward   S[own] = b (+ update vectors imported from lower clusters); per front, in elimination order, the right-looking
//             update S[struct] += (-E) S[own of the front]; at the end y = S[own] goes to Z[n + row] and the boundary
//             rows -- the update vector of the cluster's root -- to the U region, where the launches above pull it.
//   backward  S[own] = y, S[boundary] = x of the ancestors; per front, in reverse: S[own] = [F11^-1 | -G] [S[own]; S[struct]];
//             at the end x goes to the canonical state (and to Z[row] when a lower cluster gathers it).
// The update vectors of the fronts inside a cluster therefore never exist in global memory, and a whole subtree costs one
// launch per sweep instead of one per level.  The host compiles every (cluster, direction) into a program: a table of
// operations (M x K products on the FP64 tensor cores, mma.m8n8k4), 16-bit tables of resident-row indices for the K
// gathered rows and the M output rows, and the A fragments of all operations packed in consumption order, which a
// producer warp streams through a ring of 16 KB chunks (cp.async.bulk + mbarrier) starting BEFORE the grid dependency is
// resolved (factor values do not depend on the previous kernel).  A unit = 8 output rows x 16 trajectories (an even- and
// an odd-column DMMA per k-step, C fragments = 4 consecutive trajectories per lane); warp w owns units w, w+8, ... of
// an operation (up to 4), all in its half of the trajectories, so one 16-byte B load feeds all of them.
// ----------------------------------------------------------------------------------------------
constexpr int CL_NW = 16;         // consumer warps
constexpr int CL_W = 32;          // trajectories per CTA
constexpr int CL_XS = CL_W + 4;   // row stride of the resident vector (doubles)
constexpr int CL_RMAX = 2;        // units per warp and operation -> at most 128 output rows per operation
constexpr int CL_MAXROWS = 8 * CL_RMAX * CL_NW / 2;
constexpr int CL_CHUNK = 64;      // A fragments (256 bytes each) per ring chunk
constexpr int CL_STAGES = 3;
constexpr int CL_HDR = 32;        // ints: nops, nown, mroot, nimp, ustore, vfrag0, off_grow, off_ops, off_tab16, off_aux, words,
                                  //       write_z, imports: first entry of each consumer warp [12..28], off_gdof [29], store_y [30]
constexpr int CL_OPREC = 8;       // ints: M, k-steps, row blocks, mode, krow table, orow table / first output row, k parts, k-steps per chunk
constexpr int CL_RING_BYTES = CL_STAGES * CL_CHUNK * 256;
constexpr int CL_RED_BYTES = 12 * 1024;  // k-split partial sums: (parts - 1) * units KB, units * parts <= CL_NW

struct ClusterArgs {
    const int* prog;       // programs of every (cluster, direction)
    const int4* prog_loc;  // [clusters of this launch] (word offset, words, first own row if the own rows are one run else -1, own rows)
    const double* vals;    // A fragments
    double* Z;
    double* xout;
    int* diverged;
    int n, Nv, ldb, nslab, backward, prog_max_words, srow_max;
    unsigned long long* dbg;  // FCB_CLUSTER_DEBUG: 8 globaltimer stamps per CTA
};

__device__ __forceinline__ void cl_bar_consumers() { asm volatile("bar.sync 1, %0;" ::"n"(32 * CL_NW) : "memory"); }
// one 256-byte row of the resident vector -> global memory, asynchronously (bulk copy engine)
__device__ __forceinline__ void cl_row_out(double* dst, const double* src) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 256;" ::"l"(dst), "r"(smem_u32(src)) : "memory");
}

// grid = clusters of the tier x slabs of 32 trajectories (slab fastest: the CTAs of one cluster run together and share
// its factor through L2); block = (32, CL_NW + 1)
__global__ void __launch_bounds__(32 * (CL_NW + 1), 1) k_cluster_sweep(const ClusterArgs a) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x, wid = threadIdx.y;
    const int cl = blockIdx.x / a.nslab, slab0 = (blockIdx.x - cl * a.nslab) * CL_W;
    double* S = reinterpret_cast<double*>(smem);
    unsigned char* ring = smem + (size_t)a.srow_max * CL_XS * 8;
    double* red = reinterpret_cast<double*>(ring + CL_RING_BYTES);
    int* prog = reinterpret_cast<int*>(ring + CL_RING_BYTES + CL_RED_BYTES);
    const uint32_t bar_prog = smem_u32(prog) + (uint32_t)a.prog_max_words * 4u;
    const uint32_t bar_full = bar_prog + 8, bar_empty = bar_full + 8 * CL_STAGES;
    const int4 loc = __ldg(a.prog_loc + cl);
    unsigned long long* dbg = a.dbg ? a.dbg + (size_t)blockIdx.x * 8 : nullptr;
    if (dbg && lane == 0 && wid == 0) dbg[0] = globaltimer_ns();
    if (lane == 0 && wid == 0) {
        mbar_init(bar_prog, 1);
        for (int s = 0; s < CL_STAGES; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, CL_NW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_expect_tx(bar_prog, (uint32_t)loc.y * 4u);
        bulk_g2s(smem_u32(prog), a.prog + loc.x, (uint32_t)loc.y * 4u, bar_prog);
    }
    __syncthreads();
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    if (wid == CL_NW) {
        // ---------------- producer warp: the A fragments of every operation, in order; static data, so no grid dependency ----
        mbar_wait(bar_prog, 0);
        const int nops = prog[0];
        const int* ops = prog + prog[7];
        size_t frag = (size_t)(unsigned)prog[5];
        RingPos rp{0, 1};
        for (int o = 0; o < nops; ++o) {
            const int nks = ops[o * CL_OPREC + 1], nrb = ops[o * CL_OPREC + 2], cks = ops[o * CL_OPREC + 7];
            for (int k0 = 0; k0 < nks; k0 += cks) {
                const uint32_t nfr = (uint32_t)((min(nks, k0 + cks) - k0) * nrb);
                mbar_wait(bar_empty + 8 * rp.s, rp.ph);
                if (lane == 0) {
                    mbar_expect_tx(bar_full + 8 * rp.s, nfr * 256u);
                    bulk_g2s(smem_u32(ring) + rp.s * (CL_CHUNK * 256), a.vals + frag * 32, nfr * 256u, bar_full + 8 * rp.s);
                }
                frag += nfr;
                rp.advance(CL_STAGES);
            }
        }
        return;
    }

    // ---------------- consumer warps ----------------
    const int gid = lane >> 2, tig = lane & 3, half = wid & 1;
    const int tid = wid * 32 + lane;
    const size_t L = (size_t)a.ldb;
    const int nown = loc.w;
    const double* zsrc = a.Z + slab0 + (a.backward ? (size_t)a.n * L : 0);
    asm volatile("griddepcontrol.wait;" ::: "memory");  // the rows this cluster reads are complete and visible
    // resident vector: every row segment (16 bytes per thread, 16 threads per row) goes straight from global to shared
    // memory (cp.async), all in flight at once; a cluster whose own rows are one run of solver rows (every cluster of the
    // lowest tier) starts them before its program has arrived
    if (loc.z >= 0)
        for (int idx = tid; idx < nown * 16; idx += 32 * CL_NW) {
            const int i = idx >> 4, seg = (idx & 15) * 2;
            const uint32_t dst = smem_u32(S + (size_t)i * CL_XS + seg);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(zsrc + (size_t)(loc.z + i) * L + seg) : "memory");
        }
    mbar_wait(bar_prog, 0);
    if (dbg && tid == 0) dbg[1] = globaltimer_ns();  // program arrived
    const int nops = prog[0];
    const int* ops = prog + prog[7];
    const int mroot = prog[2];
    const int* grow = prog + prog[6];
    const unsigned short* tab16 = reinterpret_cast<const unsigned short*>(prog + prog[8]);
    const int* aux = prog + prog[9];
    if (loc.z < 0)
        for (int idx = tid; idx < nown * 16; idx += 32 * CL_NW) {
            const int i = idx >> 4, seg = (idx & 15) * 2;
            const uint32_t dst = smem_u32(S + (size_t)i * CL_XS + seg);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(zsrc + (size_t)grow[i] * L + seg) : "memory");
        }
    if (a.backward) {
        for (int idx = tid; idx < mroot * 16; idx += 32 * CL_NW) {
            const int j = idx >> 4, seg = (idx & 15) * 2;
            const uint32_t dst = smem_u32(S + (size_t)(nown + j) * CL_XS + seg);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(a.Z + slab0 + (size_t)aux[j] * L + seg) : "memory");
        }
    } else {
        for (int j = wid; j < mroot; j += CL_NW) S[(size_t)(nown + j) * CL_XS + lane] = 0.0;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (dbg && tid == 0) dbg[2] = globaltimer_ns();  // resident vector loaded (this thread's part)
    if (!a.backward && prog[3] > 0) {
        // update vectors of lower clusters: every resident row is summed by one warp, in the order of the list
        cl_bar_consumers();
        const double* Zc = a.Z + slab0 + lane;
        const int j1 = prog[12 + wid + 1];
        for (int j = prog[12 + wid]; j < j1; j += 8) {
            double v[8];
            int d[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (j + u < j1) {
                    v[u] = __ldcg(Zc + (size_t)aux[2 * (j + u)] * L);
                    d[u] = aux[2 * (j + u) + 1];
                }
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (j + u < j1) S[(size_t)d[u] * CL_XS + lane] += v[u];
        }
    }
    RingPos rp{0, 0};
    const unsigned char* Sb = reinterpret_cast<const unsigned char*>(S + half * 16 + 2 * gid);  // B fragments: + byte offset of the row
    for (int o = 0; o < nops; ++o) {
        const int4 q0 = *reinterpret_cast<const int4*>(ops + o * CL_OPREC);
        const int4 q1 = *reinterpret_cast<const int4*>(ops + o * CL_OPREC + 4);
        const int M = q0.x, nks = q0.y, nrb = q0.z, mode = q0.w, kp = q1.z, cks = q1.w;
        const int U = 2 * nrb;  // units (8 rows x 16 trajectories) of this operation
        // slot s = wid + CL_NW * r of this warp: unit s % U (always in this warp's half of the trajectories), k part s / U.
        // kp > 1 (few units, long K): U * kp <= CL_NW, one slot per warp, part p takes the groups of four k-steps g with
        // g % kp == p; the partial sums meet in shared memory and part 0 finishes the unit
        int un[CL_RMAX];
        bool act[CL_RMAX];
#pragma unroll
        for (int r = 0; r < CL_RMAX; ++r) {
            const int sl = wid + CL_NW * r;
            act[r] = sl < U * kp;
            un[r] = sl % U;
        }
        const int part = wid / U;  // only meaningful when kp > 1
        // byte offsets of the gathered rows, grouped [k-step / 4][lane % 4][4]: one 16-byte load = this lane's rows of 4 k-steps
        const unsigned* ko = reinterpret_cast<const unsigned*>(prog) + q1.x + tig * 4;
        cl_bar_consumers();  // the rows the previous operation wrote are in place
        double ce[CL_RMAX][2], co[CL_RMAX][2];
#pragma unroll
        for (int r = 0; r < CL_RMAX; ++r) ce[r][0] = ce[r][1] = co[r][0] = co[r][1] = 0.0;
        for (int k0 = 0; k0 < nks; k0 += cks) {
            mbar_wait(bar_full + 8 * rp.s, rp.ph);
            if (dbg && tid == 0 && o == 0 && k0 == 0) dbg[3] = globaltimer_ns();  // first chunk of A fragments arrived
            const double* ch = reinterpret_cast<const double*>(ring + rp.s * (CL_CHUNK * 256)) + lane;
            const int k1 = min(nks, k0 + cks);
            for (int ks = k0; ks < k1; ks += 4) {  // k0 and cks are multiples of 4
                if (kp > 1 && ((ks >> 2) % kp) != part) continue;  // warp-uniform
                const double* af = ch + (size_t)(ks - k0) * nrb * 32;
                if (ks + 4 <= k1) {
                    const uint4 o4 = *reinterpret_cast<const uint4*>(ko + (ks >> 2) * 16);
                    const double2 b0 = *reinterpret_cast<const double2*>(Sb + o4.x), b1 = *reinterpret_cast<const double2*>(Sb + o4.y);
                    const double2 b2 = *reinterpret_cast<const double2*>(Sb + o4.z), b3 = *reinterpret_cast<const double2*>(Sb + o4.w);
#pragma unroll
                    for (int r = 0; r < CL_RMAX; ++r)
                        if (act[r]) {  // warp-uniform
                            const double* ar = af + (un[r] >> 1) * 32;
                            const double a0 = ar[0], a1 = ar[nrb * 32], a2 = ar[2 * nrb * 32], a3 = ar[3 * nrb * 32];
                            dmma(ce[r], a0, b0.x); dmma(co[r], a0, b0.y);
                            dmma(ce[r], a1, b1.x); dmma(co[r], a1, b1.y);
                            dmma(ce[r], a2, b2.x); dmma(co[r], a2, b2.y);
                            dmma(ce[r], a3, b3.x); dmma(co[r], a3, b3.y);
                        }
                } else {
                    for (int kt = ks; kt < k1; ++kt) {
                        const double2 bq = *reinterpret_cast<const double2*>(Sb + ko[(kt >> 2) * 16 + (kt & 3)]);
#pragma unroll
                        for (int r = 0; r < CL_RMAX; ++r)
                            if (act[r]) {
                                const double av = af[(size_t)(kt - ks) * nrb * 32 + (un[r] >> 1) * 32];
                                dmma(ce[r], av, bq.x);
                                dmma(co[r], av, bq.y);
                            }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty + 8 * rp.s);
            rp.advance(CL_STAGES);
        }
        if (kp > 1) {
            // partial sums of parts 1.. -> shared memory; after the barrier part 0 adds them in order
            if (act[0] && part > 0) {
                double4* r4 = reinterpret_cast<double4*>(red) + ((size_t)(part - 1) * U + un[0]) * 32 + lane;
                *r4 = make_double4(ce[0][0], co[0][0], ce[0][1], co[0][1]);
            }
            cl_bar_consumers();  // (for a backward operation this is also "every unit has read the front's own rows")
            if (act[0] && part == 0) {
                for (int pp = 1; pp < kp; ++pp) {
                    const double4 v = reinterpret_cast<const double4*>(red)[((size_t)(pp - 1) * U + un[0]) * 32 + lane];
                    ce[0][0] += v.x; co[0][0] += v.y; ce[0][1] += v.z; co[0][1] += v.w;
                }
            } else {
                act[0] = false;
            }
        } else if (mode == 1) {
            cl_bar_consumers();  // backward: the own rows are overwritten once every unit has read them
        }
        if (mode == 0) {
            // forward: S[struct rows] += acc (the factor block is stored negated); units own disjoint (row, trajectory) tiles
            const unsigned short* orow = tab16 + q1.y;
#pragma unroll
            for (int r = 0; r < CL_RMAX; ++r)
                if (act[r]) {
                    const unsigned row = orow[(un[r] >> 1) * 8 + gid];
                    if (row != 0xffffu) {
                        double4* pq = reinterpret_cast<double4*>(S + (size_t)row * CL_XS + half * 16 + 4 * tig);
                        double4 v = *pq;
                        v.x += ce[r][0]; v.y += co[r][0]; v.z += ce[r][1]; v.w += co[r][1];
                        *pq = v;
                    }
                }
        } else {
#pragma unroll
            for (int r = 0; r < CL_RMAX; ++r) {
                const int rr = (un[r] >> 1) * 8 + gid;
                if (act[r] && rr < M)
                    *reinterpret_cast<double4*>(S + (size_t)(q1.y + rr) * CL_XS + half * 16 + 4 * tig) =
                        make_double4(ce[r][0], co[r][0], ce[r][1], co[r][1]);
            }
        }
    }
    // ---- results leave as 256-byte rows through the bulk copy engine, one row per thread and instruction
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // this thread's writes to S, before the async proxy reads them
    cl_bar_consumers();
    if (dbg && tid == 0) { dbg[4] = globaltimer_ns(); dbg[6] = (unsigned long long)nops; }  // operations done
    if (!a.backward) {
        if (prog[30]) {
            double* dst = a.Z + slab0 + (size_t)a.n * L;
            for (int i = tid; i < nown; i += 32 * CL_NW) cl_row_out(dst + (size_t)grow[i] * L, S + (size_t)i * CL_XS);
        }
        const int ustore = prog[4];
        if (ustore >= 0)
            for (int j = tid; j < mroot; j += 32 * CL_NW) cl_row_out(a.Z + slab0 + (size_t)(ustore + j) * L, S + (size_t)(nown + j) * CL_XS);
    } else {
        const int write_z = prog[11];
        const int* gdof = prog + prog[29];
        for (int i = tid; i < nown; i += 32 * CL_NW) {
            cl_row_out(a.xout + slab0 + (size_t)gdof[i] * L, S + (size_t)i * CL_XS);
            if (write_z) cl_row_out(a.Z + slab0 + (size_t)grow[i] * L, S + (size_t)i * CL_XS);
        }
        bool bad = false;  // per-trajectory flag of a non-finite velocity
        for (int i = wid; i < nown; i += CL_NW) bad = bad || (gdof[i] < a.Nv && !isfinite(S[(size_t)i * CL_XS + lane]));
        if (bad) a.diverged[slab0 + lane] = 1;
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // shared memory stays valid until the copies have read it
    if (dbg && tid == 0) dbg[5] = globaltimer_ns();
}


// ----------------------------------------------------------------------------------------------
// Matrix assembly (steady-state Newton / Picard, operator tooling): the U-dependent blocks of the linearised operator,
//   C_ab = int (U . grad phi_b) phi_a          (advection by U; the same block for both velocity components)
//   D^{ij}_ab = int phi_b (d_j U_i) phi_a      (i = row / test component, j = column / trial component)
// per cell (7-point Radon rule, exact for these degree-5 integrands) scattered into CSR value arrays of the scalar P2 pattern
// through a precomputed position map pos[cell][a][b].  Cells are coloured so that cells of one colour share no P2 node:
// one launch per colour, plain read-modify-write, no atomics, bit-reproducible.  One warp = one cell x 32 base-flow
// candidates (trajectory innermost like everything else), so a Newton step for an ensemble of base flows (Re-continuation,
// several actuation levels) is assembled in one pass.  grid = (ceil(cells of the colour / 4), ldb / 32), block = (32, 4)
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_assemble_advection(int c0, int ncell, const int* __restrict__ colour_cells,
                                                           const int* __restrict__ cell_nodes, const double* __restrict__ Jinv,
                                                           const double* __restrict__ detJ, const int* __restrict__ pos,
                                                           const double* __restrict__ U, double* __restrict__ C, double* __restrict__ D,
                                                           int nN, size_t nnz, int ldb) {
    const int ci = blockIdx.x * blockDim.y + threadIdx.y;
    if (ci >= ncell) return;
    const int e = __ldg(colour_cells + c0 + ci);
    const int b = blockIdx.y * 32 + threadIdx.x;
    const size_t L = (size_t)ldb;
    double ux[6], uy[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const int nd = __ldg(cell_nodes + (size_t)e * 6 + i);
        ux[i] = U[(size_t)nd * L + b];
        uy[i] = U[(size_t)(nd + nN) * L + b];
    }
    const double g00 = __ldg(Jinv + (size_t)e * 4), g01 = __ldg(Jinv + (size_t)e * 4 + 1), g10 = __ldg(Jinv + (size_t)e * 4 + 2),
                 g11 = __ldg(Jinv + (size_t)e * 4 + 3), det = __ldg(detJ + e);
    double ud0[7], ud1[7], G00[7], G01[7], G10[7], G11[7];
#pragma unroll
    for (int q = 0; q < 7; ++q) {
        double vx = 0.0, vy = 0.0, ax0 = 0.0, ax1 = 0.0, ay0 = 0.0, ay1 = 0.0;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            vx = fma(c_phi[q][i], ux[i], vx);
            vy = fma(c_phi[q][i], uy[i], vy);
            ax0 = fma(c_dphi[q][i][0], ux[i], ax0);
            ax1 = fma(c_dphi[q][i][1], ux[i], ax1);
            ay0 = fma(c_dphi[q][i][0], uy[i], ay0);
            ay1 = fma(c_dphi[q][i][1], uy[i], ay1);
        }
        const double wq = c_w[q] * det;
        ud0[q] = wq * (vx * g00 + vy * g01);  // wq U . grad phi_b = dphi_ref[b][0] ud0 + dphi_ref[b][1] ud1
        ud1[q] = wq * (vx * g10 + vy * g11);
        G00[q] = wq * (ax0 * g00 + ax1 * g10);  // d_x U_x
        G01[q] = wq * (ax0 * g01 + ax1 * g11);  // d_y U_x
        G10[q] = wq * (ay0 * g00 + ay1 * g10);  // d_x U_y
        G11[q] = wq * (ay0 * g01 + ay1 * g11);  // d_y U_y
    }
    const int* pe = pos + (size_t)e * 36;
    for (int a = 0; a < 6; ++a) {
#pragma unroll
        for (int bb = 0; bb < 6; ++bb) {
            double cab = 0.0, dxx = 0.0, dxy = 0.0, dyx = 0.0, dyy = 0.0;
#pragma unroll
            for (int q = 0; q < 7; ++q) {
                const double pa = c_phi[q][a];
                cab = fma(pa, fma(c_dphi[q][bb][0], ud0[q], c_dphi[q][bb][1] * ud1[q]), cab);
                const double pab = pa * c_phi[q][bb];
                dxx = fma(pab, G00[q], dxx);
                dxy = fma(pab, G01[q], dxy);
                dyx = fma(pab, G10[q], dyx);
                dyy = fma(pab, G11[q], dyy);
            }
            const size_t p0 = (size_t)__ldg(pe + a * 6 + bb) * L + b;
            C[p0] += cab;
            if (D) {
                D[p0] += dxx;
                D[nnz * L + p0] += dxy;
                D[2 * nnz * L + p0] += dyx;
                D[3 * nnz * L + p0] += dyy;
            }
        }
    }
}


// ----------------------------------------------------------------------------------------------
// Numeric multifrontal factorisation (setup): fronts are dense (w+m) x (w+m) matrices, row-major, leading dimension w+m,
// all in one workspace; a level of the elimination tree (fronts of equal height) is one launch of each kernel.
//   k_ff_assemble  F = entries of the sparse matrix + Schur complements of the children (extend-add, children in order)
//   k_ff_invert    F11 <- F11^-1 in place (Gauss-Jordan, partial pivoting), copy to the output, growth estimate
//   k_ff_gemm      mode 0: E = F21 F11^-1   mode 1: G = F11^-1 F12   mode 2: F22 -= E F12   (32 x 32 tiles, FP64 FMA)
// ----------------------------------------------------------------------------------------------
struct FrontDesc {
    long long foff, eoff, ioff;  // workspace / E,G output / F11^-1 output offsets (doubles)
    int w, m;
};

// grid = fronts of the level, block = 512
__global__ void __launch_bounds__(512) k_ff_assemble(const int* __restrict__ fronts, const FrontDesc* __restrict__ fd, const long long* __restrict__ a_ptr,
                                                    const int* __restrict__ a_src, const int* __restrict__ a_dst, const double* __restrict__ avals,
                                                    const long long* __restrict__ c_ptr, const int* __restrict__ c_front,
                                                    const long long* __restrict__ c_lptr, const int* __restrict__ c_loc, double* __restrict__ W) {
    const int f = fronts[blockIdx.x];
    const FrontDesc d = fd[f];
    const int s = d.w + d.m;
    double* F = W + d.foff;
    for (long long k = a_ptr[f] + threadIdx.x; k < a_ptr[f + 1]; k += blockDim.x) F[a_dst[k]] = avals[a_src[k]];
    for (long long c = c_ptr[f]; c < c_ptr[f + 1]; ++c) {
        __syncthreads();  // children one after the other: a fixed summation order for the entries two children share
        const FrontDesc dc = fd[c_front[c]];
        const int sc = dc.w + dc.m, mc = dc.m;
        const double* CB = W + dc.foff + (size_t)dc.w * sc + dc.w;  // Schur complement of the child (its F22 after elimination)
        const int* loc = c_loc + c_lptr[c];
        for (long long t = threadIdx.x; t < (long long)mc * mc; t += blockDim.x) {
            const int i = (int)(t / mc), j = (int)(t - (long long)i * mc);
            F[(size_t)loc[i] * s + loc[j]] += CB[(size_t)i * sc + j];
        }
    }
}

// grid = fronts of the level, block = 512; dynamic shared memory: w ints (pivot rows)
__global__ void __launch_bounds__(512) k_ff_invert(const int* __restrict__ fronts, const FrontDesc* __restrict__ fd, double* __restrict__ W,
                                                  double* __restrict__ Finv, double* __restrict__ growth) {
    extern __shared__ int piv[];
    __shared__ double red_v[16];
    __shared__ int red_i[16];
    __shared__ double s_scale;
    const int f = fronts[blockIdx.x];
    const FrontDesc d = fd[f];
    const int w = d.w, s = d.w + d.m, tid = threadIdx.x, nt = blockDim.x;
    double* A = W + d.foff;  // F11 = rows / columns [0, w), leading dimension s
    // max |F11| (growth estimate)
    double mx = 0.0;
    for (int t = tid; t < w * w; t += nt) mx = fmax(mx, fabs(A[(size_t)(t / w) * s + (t % w)]));
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((tid & 31) == 0) red_v[tid >> 5] = mx;
    __syncthreads();
    if (tid == 0) {
        double v = 0.0;
        for (int k = 0; k < nt / 32; ++k) v = fmax(v, red_v[k]);
        s_scale = v;
    }
    __syncthreads();
    for (int k = 0; k < w; ++k) {
        // pivot: largest |A[i][k]|, i >= k (ties: the smallest row, deterministic)
        double bv = -1.0;
        int bi = k;
        for (int i = k + tid; i < w; i += nt) {
            const double v = fabs(A[(size_t)i * s + k]);
            if (v > bv) { bv = v; bi = i; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if ((tid & 31) == 0) { red_v[tid >> 5] = bv; red_i[tid >> 5] = bi; }
        __syncthreads();
        if (tid == 0) {
            double v = red_v[0];
            int ix = red_i[0];
            for (int q = 1; q < nt / 32; ++q)
                if (red_v[q] > v || (red_v[q] == v && red_i[q] < ix)) { v = red_v[q]; ix = red_i[q]; }
            piv[k] = ix;
        }
        __syncthreads();
        const int p = piv[k];
        if (p != k)
            for (int j = tid; j < w; j += nt) {
                const double t0 = A[(size_t)k * s + j];
                A[(size_t)k * s + j] = A[(size_t)p * s + j];
                A[(size_t)p * s + j] = t0;
            }
        __syncthreads();
        const double pinv = 1.0 / A[(size_t)k * s + k];
        __syncthreads();
        // row k: the pivot column becomes the unit vector's image
        for (int j = tid; j < w; j += nt) A[(size_t)k * s + j] = (j == k ? 1.0 : A[(size_t)k * s + j]) * pinv;
        __syncthreads();
        // every other row i: A[i][:] -= A[i][k] * row k, with A[i][k] replaced by 0 first
        for (int i = tid >> 5; i < w; i += nt >> 5) {
            if (i == k) continue;
            const double fk = A[(size_t)i * s + k];
            __syncwarp();
            for (int j = tid & 31; j < w; j += 32) {
                const double rk = A[(size_t)k * s + j];
                const double v = (j == k ? 0.0 : A[(size_t)i * s + j]);
                A[(size_t)i * s + j] = fma(-fk, rk, v);
            }
        }
        __syncthreads();
    }
    // undo the row exchanges as column exchanges, last first
    for (int k = w - 1; k >= 0; --k) {
        const int p = piv[k];
        if (p != k)
            for (int i = tid; i < w; i += nt) {
                const double t0 = A[(size_t)i * s + k];
                A[(size_t)i * s + k] = A[(size_t)i * s + p];
                A[(size_t)i * s + p] = t0;
            }
        __syncthreads();
    }
    double mi = 0.0;
    double* out = Finv + d.ioff;
    for (int t = tid; t < w * w; t += nt) {
        const double v = A[(size_t)(t / w) * s + (t % w)];
        out[t] = v;
        mi = isfinite(v) ? fmax(mi, fabs(v)) : INFINITY;  // a NaN / Inf anywhere flags the front as singular
    }
    for (int o = 16; o > 0; o >>= 1) mi = fmax(mi, __shfl_xor_sync(0xffffffffu, mi, o));
    __syncthreads();
    if ((tid & 31) == 0) red_v[tid >> 5] = mi;
    __syncthreads();
    if (tid == 0) {
        double v = 0.0;
        for (int k = 0; k < nt / 32; ++k) v = fmax(v, red_v[k]);
        growth[f] = (w > 0) ? (isfinite(v) ? v * s_scale : INFINITY) : 0.0;
    }
}

// grid = (fronts of the level, tiles); block = (16, 16), each thread 2 x 2 outputs of a 32 x 32 tile
__global__ void __launch_bounds__(256) k_ff_gemm(int mode, const int* __restrict__ fronts, const FrontDesc* __restrict__ fd, double* __restrict__ W,
                                                double* __restrict__ Eo, double* __restrict__ Go) {
    __shared__ double As[32][33], Bs[32][33];
    const int f = fronts[blockIdx.x];
    const FrontDesc d = fd[f];
    const int w = d.w, m = d.m, s = w + m;
    double* F = W + d.foff;
    // C [M x N] = (or -=) A [M x K] B [K x N]
    int M, N, K, lda, ldb, ldc;
    const double *A, *B;
    double* Cc;
    if (mode == 0)      { M = m; N = w; K = w; A = F + (size_t)w * s; lda = s; B = F; ldb = s; Cc = Eo + d.eoff; ldc = w; }
    else if (mode == 1) { M = w; N = m; K = w; A = F; lda = s; B = F + w; ldb = s; Cc = Go + d.eoff; ldc = m; }
    else                { M = m; N = m; K = w; A = Eo + d.eoff; lda = w; B = F + w; ldb = s; Cc = F + (size_t)w * s + w; ldc = s; }
    const int tn = (N + 31) / 32, tm = (M + 31) / 32;
    for (int tile = blockIdx.y; tile < tm * tn; tile += gridDim.y) {
        const int r0 = (tile / tn) * 32, c0 = (tile % tn) * 32;
        const int tx = threadIdx.x, ty = threadIdx.y;
        double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
        for (int k0 = 0; k0 < K; k0 += 32) {
            for (int q = ty; q < 32; q += 16)
                for (int r = tx; r < 32; r += 16) {
                    As[q][r] = (r0 + q < M && k0 + r < K) ? A[(size_t)(r0 + q) * lda + k0 + r] : 0.0;
                    Bs[q][r] = (k0 + q < K && c0 + r < N) ? B[(size_t)(k0 + q) * ldb + c0 + r] : 0.0;
                }
            __syncthreads();
#pragma unroll 8
            for (int kk = 0; kk < 32; ++kk) {
                const double a0 = As[ty][kk], a1 = As[ty + 16][kk], b0 = Bs[kk][tx], b1 = Bs[kk][tx + 16];
                acc[0][0] = fma(a0, b0, acc[0][0]); acc[0][1] = fma(a0, b1, acc[0][1]);
                acc[1][0] = fma(a1, b0, acc[1][0]); acc[1][1] = fma(a1, b1, acc[1][1]);
            }
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int r = r0 + ty + 16 * i, c = c0 + tx + 16 * j;
                if (r < M && c < N) {
                    double* o = Cc + (size_t)r * ldc + c;
                    *o = (mode == 2) ? *o - acc[i][j] : acc[i][j];
                }
            }
    }
}

// sensors + energy.  grid = ldb/32, block = (32, MEAS_WARPS)
constexpr int MEAS_WARPS = 32;
__global__ void __launch_bounds__(32 * MEAS_WARPS) k_measure(int ns, const int* __restrict__ sptr, const int* __restrict__ sidx,
                                                           const double* __restrict__ sval, const double* __restrict__ up,
                                                           double* __restrict__ y, const double* __restrict__ epart, int nblk,
                                                           double* __restrict__ dE, int na, const double* __restrict__ uctrl,
                                                           double* __restrict__ uctrl_prev, int ldb) {
    const int b = blockIdx.x * 32 + threadIdx.x;
    const int ty = threadIdx.y;
    // Crank-Nicolson: this step's control becomes "the previous step's" (the averaged body force of the next rhs); a
    // memcpy node in the step graph for these few kB cost 70 us
    if (uctrl_prev)
        for (int k = ty; k < na; k += MEAS_WARPS) uctrl_prev[(size_t)k * ldb + b] = uctrl[(size_t)k * ldb + b];
    for (int s = ty; s < ns; s += MEAS_WARPS) {
        double acc = 0.0;
        for (int j = __ldg(sptr + s); j < __ldg(sptr + s + 1); ++j)
            acc = fma(__ldg(sval + j), up[(size_t)__ldg(sidx + j) * ldb + b], acc);
        y[(size_t)s * ldb + b] = acc;
    }
    // energy: every warp sums a contiguous chunk of the patch partials (loads independent of each other), then the
    // warps' sums are combined in a fixed order (deterministic)
    const int chunk = (nblk + MEAS_WARPS - 1) / MEAS_WARPS;
    const int k0 = ty * chunk, k1 = min(nblk, k0 + chunk);
    double e0 = 0.0, e1 = 0.0, e2 = 0.0, e3 = 0.0;
    int k = k0;
    for (; k + 3 < k1; k += 4) {
        e0 += epart[(size_t)k * ldb + b];
        e1 += epart[(size_t)(k + 1) * ldb + b];
        e2 += epart[(size_t)(k + 2) * ldb + b];
        e3 += epart[(size_t)(k + 3) * ldb + b];
    }
    for (; k < k1; ++k) e0 += epart[(size_t)k * ldb + b];
    __shared__ double se[MEAS_WARPS][32];
    se[ty][threadIdx.x] = (e0 + e1) + (e2 + e3);
    __syncthreads();
    if (ty == 0) {
        double s = 0.0;
        for (int w = 0; w < MEAS_WARPS; ++w) s += se[w][threadIdx.x];
        dE[b] = 0.5 * s;
    }
}

// One LTI controller per trajectory (controller.py:157-158): uses the PRE-update state for the output.
// grid = ldb/32, block = (32, CTRL_ROWS): thread (b, i) owns row i of the state update of trajectory b, rows < nu also an
// output; the rows of one trajectory meet in shared memory for the actuator fan-out.
constexpr int CTRL_ROWS = 16;
__global__ void __launch_bounds__(32 * CTRL_ROWS) k_controller(int nx, int ny, int nu, int ns, int na, const double* __restrict__ Ad,
                                                              const double* __restrict__ Bd, const double* __restrict__ Cd,
                                                              const double* __restrict__ Dd, const double* __restrict__ Ky,
                                                              const double* __restrict__ Fu, const double* __restrict__ y,
                                                              const double* __restrict__ xin, double* __restrict__ xout,
                                                              double* __restrict__ uctrl, int ldb) {
    const int b = blockIdx.x * 32 + threadIdx.x;
    const int r = threadIdx.y;
    const size_t L = (size_t)ldb;
    double v[8];
    for (int i = 0; i < ny; ++i) {
        double s = 0.0;
        for (int j = 0; j < ns; ++j) s = fma(__ldg(Ky + i * ns + j), y[j * L + b], s);
        v[i] = s;
    }
    __shared__ double uo[8][32];
    for (int o = r; o < nu; o += CTRL_ROWS) {
        double s = 0.0;
        for (int j = 0; j < nx; ++j) s = fma(Cd[(size_t)(o * nx + j) * L + b], xin[j * L + b], s);
        for (int j = 0; j < ny; ++j) s = fma(Dd[(size_t)(o * ny + j) * L + b], v[j], s);
        uo[o][threadIdx.x] = s;
    }
    for (int i = r; i < nx; i += CTRL_ROWS) {
        double s = 0.0;
        for (int j = 0; j < nx; ++j) s = fma(Ad[(size_t)(i * nx + j) * L + b], xin[j * L + b], s);
        for (int j = 0; j < ny; ++j) s = fma(Bd[(size_t)(i * ny + j) * L + b], v[j], s);
        xout[i * L + b] = s;
    }
    __syncthreads();
    for (int a = r; a < na; a += CTRL_ROWS) {
        double s = 0.0;
        for (int o = 0; o < nu; ++o) s = fma(__ldg(Fu + a * nu + o), uo[o][threadIdx.x], s);
        uctrl[a * L + b] = s;
    }
}

// Run control block of the device-resident loops (one per handle, in device memory): the graphs of a step read the series
// pointer, its capacity and the step counter from here, so a run of any length replays the same graphs chunk by chunk.
struct RunCtl {
    double* series;         // [capacity][ncol][ldb] chunk buffer, or capacity == 0: logging off
    const double* useries;  // [capacity][na][ldb] open-loop control inputs of the chunk
    int capacity;
    int step;               // step inside the current chunk
};

// open loop: this step's control inputs come from the staged chunk of the caller's u_ctrl series.  grid = ldb/32, block = 32
__global__ void k_load_ctrl(int na, const RunCtl* __restrict__ ctl, double* __restrict__ uctrl, int ldb) {
    const int b = blockIdx.x * 32 + threadIdx.x;
    const double* src = ctl->useries + (size_t)ctl->step * na * ldb;
    for (int k = 0; k < na; ++k) uctrl[(size_t)k * ldb + b] = src[(size_t)k * ldb + b];
}

// series[step][col][b], columns (dE, u_ctrl_1..na, y_1..ns); single CTA, then the step counter ticks.
// also accumulates, per trajectory and in time order, the sums the reference's cost functions are made of
// (utils/optim.py:231-288): costs[0] += dE, costs[1] += sum_k u_ctrl_k^2, costs[2] = dE of the last step
__global__ void k_log(int na, int ns, const double* __restrict__ dE, const double* __restrict__ uctrl,
                      const double* __restrict__ y, RunCtl* __restrict__ ctl, double* __restrict__ costs, int ldb) {
    const int step = ctl->step;
    const int ncol = 1 + na + ns;
    for (int b = threadIdx.x; b < ldb; b += blockDim.x) {
        double u2 = 0.0;
        for (int k = 0; k < na; ++k) u2 = fma(uctrl[(size_t)k * ldb + b], uctrl[(size_t)k * ldb + b], u2);
        const double e = dE[b];
        costs[b] += e;
        costs[ldb + b] += u2;
        costs[2 * ldb + b] = e;
    }
    if (step < ctl->capacity) {
        double* row = ctl->series + (size_t)step * ncol * ldb;
        for (int i = threadIdx.x; i < ncol * ldb; i += blockDim.x) {
            const int col = i / ldb, b = i - col * ldb;
            double v;
            if (col == 0) v = dE[b];
            else if (col <= na) v = uctrl[(size_t)(col - 1) * ldb + b];
            else v = y[(size_t)(col - 1 - na) * ldb + b];
            row[i] = v;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) ctl->step = step + 1;
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
struct DevPlan {
    int n = 0, nU = 0, njobs = 0, nlaunch = 0;
    int *srec = nullptr, *jrec = nullptr, *cta_sptr = nullptr, *cta_jptr = nullptr;
    int asm_n = 0, *asm_ptr = nullptr, *asm_src = nullptr, *asm_dst = nullptr;
    std::vector<int> asm_lptr;  // [nlaunch+1]: gather-sum rows [asm_lptr[l], asm_lptr[l+1]) run right before launch l
    double* vals = nullptr;
    struct Launch { int grid, nwc, nslab, nstages, slots, cta_off, ksplit; };
    std::vector<Launch> launches;
    int n_forward = 0;
    long long nstages_total = 0, packed_doubles = 0;
    // subtree clusters (k_cluster_sweep)
    struct Tier { int first, count; };
    std::vector<Tier> tiers;
    int ncluster = 0, cl_srow_max = 0, cl_prog_max = 0, cl_smem = 0;
    int* cprog = nullptr;
    int4* cprog_loc[2] = {nullptr, nullptr};  // forward / backward: (word offset, words, first own row or -1, own rows) per cluster
    double* cvals = nullptr;
    long long cl_frags = 0;
};

}  // namespace

struct fcb_context {
    int device = 0, num_sms = 0, smem_per_sm = 0, force_nrb = 0, force_nwc = 0, max_nwc = 4, allow_ksplit = 1;
    SweepMaps zmaps[4];  // TMA descriptors of Z for CTA widths of 32, 64, 128, 256 trajectories
    double want_ctas_per_sm = 2.0;          // a launch narrows its CTAs / shortens its tiles until it has this many CTAs per SM
    double want_ctas_fwd = 1.0;             // the same for the forward launches (FCB_SWEEP_WANT_FWD): their three-plane gathers favour wide CTAs
                                            // and tall tiles over more CTAs (measured on B200, cylinder x 256: forward 0.282 -> 0.271 ms for 0.5..1.5)
    int kslots = 24;                        // ... and for k-split CTAs, whose 4 warps each need stages in flight (FCB_SWEEP_KSLOTS)
    int force_slots[4] = {48, 36, 12, 12};  // gathered rows per ring stage for those widths (FCB_SWEEP_SLOTS=a,b,c,d)
    unsigned long long* sweep_dbg = nullptr;  // FCB_SWEEP_DEBUG=<file>: per-CTA timeline of the sweeps of a profiled step
    unsigned long long* cl_dbg = nullptr;     // FCB_CLUSTER_DEBUG=<file>: per-CTA timeline of the cluster sweeps (8 tiers x 2 directions x 8192 CTAs)
    cudaStream_t stream = nullptr;
    std::string error;
    int B = 0, ldb = 0;
    int nT = 0, nN = 0, nV = 0, Nv = 0, N = 0, n = 0, nbc = 0, na = 0, ns = 0;
    double dt = 0.0;
    int nonlinear = 1;
    // constant device data
    int *cell_nodes = nullptr, *perm = nullptr, *iperm = nullptr;
    double *Jinv = nullptr, *detJ = nullptr, *bc_shape = nullptr, *ctrl_rhs[2] = {nullptr, nullptr};
    int *sensor_ptr = nullptr, *sensor_idx = nullptr;
    double* sensor_val = nullptr;
    int nblk_total = 0;  // rows of the energy partial sums (one per element patch)
    // patch form of the element kernel
    int use_pdl = 1;
    int npatch = 0, nshared = 0, patch_smem = 0, patch_tab_off = 0;
    int *pcell_ptr = nullptr, *pcnode = nullptr;
    double* pgeo = nullptr;
    int *pnode_ptr = nullptr, *pnode_dst = nullptr, *mptr = nullptr, *msrc = nullptr,
        *mnode = nullptr, *prow = nullptr, *mrow = nullptr, *pacc_rows = nullptr;
    unsigned char *plnode = nullptr, *psrc = nullptr;
    double* pscratch = nullptr;
    DevPlan plan[2];
    // state
    double *up[2] = {nullptr, nullptr}, *avec = nullptr, *bvec[2] = {nullptr, nullptr}, *Z = nullptr;
    double *epart = nullptr, *uctrl = nullptr, *y = nullptr, *dE = nullptr;
    int* diverged = nullptr;
    int parity = 0, order = 1;
    bool rhs_ready = false;  // Z[0,n) holds the fused right-hand side of the next step
    int scheme = 0;          // 0 = BDF1 -> BDF2, 1 = Crank-Nicolson
    int *cn_ptr = nullptr, *cn_idx = nullptr;
    double *cn_val = nullptr, *uctrl_prev = nullptr, *ccoef_prev = nullptr;
    int ncrow_prev = 0;
    int* crow_prev = nullptr;
    // packed form of the same operator for k_spmm_mma (build_spmm)
    int sp_nblk = 0, sp_smem = 0;
    bool sp_csr = false;  // FCB_SPMM_CSR=1: run the scalar CSR kernel instead (A/B measurements)
    int *sp_ucols = nullptr, *sp_ginfo = nullptr;
    unsigned short* sp_kslots = nullptr;
    double* sp_avals = nullptr;
    int ncrow = 0;           // solver rows with a non-zero control coefficient (BDF2), their coefficients [na, ncrow]
    int* crow = nullptr;
    double* ccoef = nullptr;
    int* bc_dofs = nullptr;
    bool have_state = false;
    // controllers
    bool have_ctrl = false;
    int nx = 0, ny = 0, nu = 0;
    double *Ad = nullptr, *Bd = nullptr, *Cd = nullptr, *Dd = nullptr, *Ky = nullptr, *Fu = nullptr;
    double* xk[2] = {nullptr, nullptr};
    int xparity = 0;
    double* series = nullptr;   // chunk buffer of the device-resident loops [series_capacity][ncol][ldb]
    double* useries = nullptr;  // open loop: staged control inputs of a chunk [series_capacity][na][ldb]
    int series_capacity = 0;
    int series_chunk = 1024;    // steps per chunk (FCB_SERIES_CHUNK): long runs stream their series out chunk by chunk
    RunCtl* ctl = nullptr;
    double* costs = nullptr;  // [3][ldb] running sums of the closed-loop run (k_log)
    double *pin_in = nullptr, *pin_out = nullptr;  // pinned staging of fcb_step's host arguments
    // graphs: [parity] for a BDF2 step, [parity][xparity] for a closed-loop BDF2 step
    cudaGraphExec_t g_step[2] = {nullptr, nullptr};
    int g_step_nodes[2] = {0, 0};
    cudaGraphExec_t g_loop[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    int g_loop_nodes[2][2] = {{0, 0}, {0, 0}};
    cudaGraphExec_t g_open[2] = {nullptr, nullptr};  // open-loop run: load u_ctrl -> step -> log
    int g_open_nodes[2] = {0, 0};
    long long launches = 0;
    int phase_launches[FCB_NPHASES] = {0};
    cudaEvent_t ev[FCB_NPHASES + 1] = {nullptr};
};

namespace {

int fail(fcb_context* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->error = buf;
    else g_create_error = buf;
    return code;
}

#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess)                                                                            \
            return fail(h, FCB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

template <typename T>
int upload(fcb_context* h, T** dst, const T* src, size_t count) {
    *dst = nullptr;
    if (count == 0) return FCB_OK;
    CK(cudaMalloc((void**)dst, count * sizeof(T)));
    if (src) CK(cudaMemcpyAsync(*dst, src, count * sizeof(T), cudaMemcpyDefault, h->stream));
    else CK(cudaMemsetAsync(*dst, 0, count * sizeof(T), h->stream));
    return FCB_OK;
}

#define TRY(expr)                  \
    do {                           \
        int rc_ = (expr);          \
        if (rc_ != FCB_OK) return rc_; \
    } while (0)

// Compile a SolvePlan into per-CTA instruction streams (see k_front_sweep) and upload it.
// Tiling policy per launch (the blocks of one elimination-tree level): tiles as tall as possible (up to
// 4 row blocks of 8) while the launch still has ~4 warp-jobs per SM; CTAs as wide as possible (up to 8
// warps x 32 trajectories, sharing one copy of V and of the gathered rows) while there is at least
// one CTA per SM.
int upload_plan(fcb_context* h, DevPlan& d, const fcb_plan& p, const int32_t* perm) {
    d.n = p.n;
    d.nU = p.nU;
    d.nlaunch = p.nlaunch;
    d.n_forward = p.n_forward_launches;
    const int zrow = 2 * p.n + p.nU;
    for (int i = 0; i < p.nblocks; ++i) {
        const int K = p.blk_K[i], M = p.blk_M[i], nsrc = p.blk_nsrc[i];
        if (K < 0 || M < 0 || (nsrc != 1 && nsrc != 3) || p.blk_out0[i] < 0 || p.blk_out0[i] + M > zrow ||
            (p.blk_ystore[i] >= 0 && p.blk_ystore[i] + K > zrow) || p.blk_iptr[i] < 0 || p.blk_vptr[i] < 0)
            return fail(h, FCB_ERR_INVALID, "plan block %d is malformed", i);
        for (long long k = p.blk_iptr[i]; k < p.blk_iptr[i] + K; ++k)
            if (p.i0[k] < 0 || p.i0[k] > zrow || (nsrc == 3 && (p.i1[k] < 0 || p.i1[k] > zrow || p.i2[k] < 0 || p.i2[k] > zrow)))
                return fail(h, FCB_ERR_INVALID, "plan block %d: gather index out of range", i);
        if (p.blk_eptr[i] >= 0)
            for (int r = 0; r < M; ++r)
                if (p.e0[p.blk_eptr[i] + r] >= zrow || p.e1[p.blk_eptr[i] + r] >= zrow)
                    return fail(h, FCB_ERR_INVALID, "plan block %d: seed index out of range", i);
    }
    struct Tile { int blk, r0, nr, nrb, k_lo, k_hi; bool ystore; };  // rows [r0, r0+nr) of a block; ystore rows [k_lo,k_hi)
    std::vector<int> srec, jrec, cta_sptr, cta_jptr;
    std::vector<double> packed;
    d.launches.clear();
    const int nw_all = h->ldb / 32;  // 32-trajectory warps needed to cover the ensemble
    for (int l = 0; l < p.nlaunch; ++l) {
        const int b0 = p.launch_ptr[l], b1 = p.launch_ptr[l + 1];
        DevPlan::Launch L{0, 4, 1, 4, 12, (int)cta_sptr.size(), 0};
        if (b1 <= b0) { d.launches.push_back(L); continue; }
        // ---- tile height and CTA width: tall tiles (V reuse, fewer re-reads of the gathered rows) and wide
        // CTAs (one copy of V per CTA) as long as the launch still fills the machine; narrow first, then shorten
        auto njobs_for = [&](int nrb) {
            long long c = 0;
            for (int b = b0; b < b1; ++b) c += p.blk_M[b] > 0 ? (p.blk_M[b] + 8 * nrb - 1) / (8 * nrb) : (p.blk_K[b] + 15) / 16;
            return c;
        };
        int nrb_cap = 4, nwc = std::min(h->max_nwc, nw_all);
        while (nw_all % nwc) nwc >>= 1;
        const long long want = (long long)((l < p.n_forward_launches && h->want_ctas_fwd > 0.0 ? h->want_ctas_fwd : h->want_ctas_per_sm) * h->num_sms);
        while (nwc > 1 && njobs_for(nrb_cap) * (nw_all / nwc) < want) nwc >>= 1;
        // a launch that is still too small at one 32-trajectory tile per CTA splits K over the CTA's four consumer warps
        L.ksplit = (nwc == 1 && h->allow_ksplit) ? 1 : 0;
        while (nrb_cap > 1 && njobs_for(nrb_cap) * (nw_all / nwc) < want) --nrb_cap;
        if (h->force_nrb > 0) nrb_cap = h->force_nrb;
        if (h->force_nwc > 0 && nw_all % h->force_nwc == 0) nwc = h->force_nwc;
        L.nwc = nwc;
        L.nslab = nw_all / nwc;
        std::vector<Tile> tiles;
        for (int b = b0; b < b1; ++b) {
            const int M = p.blk_M[b], K = p.blk_K[b];
            if (M == 0) {
                // store-only block (the root of the tree): rows are independent, cut into short copy jobs
                for (int k = 0; k < K; k += 16) tiles.push_back({b, 0, 0, 0, k, std::min(K, k + 16), true});
                continue;
            }
            const int nblk = (M + 7) / 8, ntile = (nblk + nrb_cap - 1) / nrb_cap;
            int r0 = 0;
            for (int t = 0; t < ntile; ++t) {
                const int nb = nblk / ntile + (t < nblk % ntile ? 1 : 0), nr = std::min(8 * nb, M - r0);
                tiles.push_back({b, r0, nr, nb, 0, K, t == 0 && p.blk_ystore[b] >= 0});
                r0 += nr;
            }
        }
        const int nj = (int)tiles.size();
        const int min_ctas = nwc == 8 ? SV_MINCTAS8 : ((nwc == 4 || L.ksplit) ? SV_MINCTAS4 : 4);
        // rows per stage: ~24 KB of gathered rows for the wide CTAs, less for narrow ones (their V slice is as large)
        L.slots = L.ksplit ? h->kslots : h->force_slots[nwc == 8 ? 3 : nwc == 4 ? 2 : nwc == 2 ? 1 : 0];
        const int stage_bytes = nwc == 8 ? SweepCfg<8>::stage_bytes(L.slots) : nwc == 4 ? SweepCfg<4>::stage_bytes(L.slots)
                                : nwc == 2 ? SweepCfg<2>::stage_bytes(L.slots) : SweepCfg<1>::stage_bytes(L.slots);
        const long long nctas = (long long)nj * L.nslab;
        const int per_sm = (nctas >= (long long)min_ctas * h->num_sms) ? min_ctas : std::max(1, (int)((nctas + h->num_sms - 1) / h->num_sms));
        L.grid = std::min(nj, std::max(1, per_sm * h->num_sms / L.nslab));
        L.nstages = std::max(2, std::min(SV_MAXSTAGES, (h->smem_per_sm / per_sm - 1024 - 2 * SV_MAXSTAGES * 8 - (L.ksplit ? SV_RED_BYTES : 0)) / stage_bytes));
        const int kc1 = L.slots & ~3, kc3 = (L.slots / 3) & ~3;
        const int stride_w = 32 * nwc;  // trajectories per CTA
        // ---- longest-processing-time assignment of tiles to CTAs (cost in rough SM cycles)
        auto ktile = [&](const Tile& t) { return t.nrb > 0 ? p.blk_K[t.blk] : t.k_hi - t.k_lo; };
        auto stages_of = [&](const Tile& t) {
            const int K4 = (ktile(t) + 3) & ~3, kc = p.blk_nsrc[t.blk] == 3 ? kc3 : kc1;
            return std::max(1, (K4 + kc - 1) / kc);
        };
        std::vector<std::pair<long long, int>> order(nj);
        for (int q = 0; q < nj; ++q)
            order[q] = {-(600LL + 150LL * stages_of(tiles[q]) + 8LL * ((ktile(tiles[q]) + 3) & ~3) * std::max(1, tiles[q].nrb)), q};
        std::sort(order.begin(), order.end());
        std::vector<std::vector<int>> mine(L.grid);
        std::vector<std::pair<long long, int>> heap(L.grid);  // (load, cta) min-heap
        for (int c = 0; c < L.grid; ++c) heap[c] = {0, c};
        auto cmp = [](const std::pair<long long, int>& a, const std::pair<long long, int>& b) { return a > b; };
        std::make_heap(heap.begin(), heap.end(), cmp);
        for (auto& [negcost, q] : order) {
            std::pop_heap(heap.begin(), heap.end(), cmp);
            heap.back().first += -negcost;
            mine[heap.back().second].push_back(q);
            std::push_heap(heap.begin(), heap.end(), cmp);
        }
        // ---- emit the streams; V goes out in MMA A-fragment order [K4/4][nrb][8 rows][4 k]
        for (int c = 0; c < L.grid; ++c) {
            cta_sptr.push_back((int)(srec.size() / SV_SREC));
            cta_jptr.push_back((int)(jrec.size() / SV_JREC));
            for (int q : mine[c]) {
                const Tile& t = tiles[q];
                const int b = t.blk, Kb = p.blk_K[b], nsrc = p.blk_nsrc[b];
                const int K = t.k_hi - t.k_lo, K4 = (K + 3) & ~3, kc = nsrc == 3 ? kc3 : kc1;
                const size_t jb = jrec.size();
                jrec.resize(jb + SV_JREC, -1);
                jrec[jb + 0] = K; jrec[jb + 1] = t.nrb; jrec[jb + 2] = t.nr; jrec[jb + 3] = nsrc;
                jrec[jb + 4] = p.blk_out0[b] + t.r0;
                jrec[jb + 5] = t.ystore ? p.blk_ystore[b] + t.k_lo : -1;
                const bool seeded = t.nrb > 0 && p.blk_eptr[b] >= 0;
                const bool is_x = t.nrb > 0 && l >= p.n_forward_launches && p.blk_out0[b] + p.blk_M[b] <= p.n;  // rows of the solution
                if (seeded && is_x) return fail(h, FCB_ERR_INVALID, "plan block %d: a backward block cannot carry seed rows", b);
                // blk_ystore == -2 on a block of solution rows: no other block gathers them, they only go out in canonical
                // numbering (seed mode 3) and the b rows they would overwrite stay intact
                if (p.blk_ystore[b] == -2 && !is_x) return fail(h, FCB_ERR_INVALID, "plan block %d: ystore -2 on a block that does not produce solution rows", b);
                jrec[jb + 6] = seeded ? 1 : (is_x ? (p.blk_ystore[b] == -2 ? 3 : 2) : 0);
                if (seeded)
                    for (int r = 0; r < t.nr; ++r) {
                        jrec[jb + 8 + r] = p.e0[p.blk_eptr[b] + t.r0 + r];
                        jrec[jb + 40 + r] = p.e1[p.blk_eptr[b] + t.r0 + r];
                    }
                if (is_x)
                    for (int r = 0; r < t.nr; ++r) jrec[jb + 8 + r] = perm[p.blk_out0[b] + t.r0 + r];
                const size_t v0 = packed.size();
                if (t.nrb > 0) {
                    packed.resize(v0 + (size_t)K4 * 8 * t.nrb, 0.0);
                    const double* V = p.vals + p.blk_vptr[b];
                    for (int r = 0; r < t.nr; ++r) {
                        const double* vr = V + (size_t)(t.r0 + r) * Kb;
                        double* dst = packed.data() + v0 + (size_t)(r >> 3) * 32 + (size_t)(r & 7) * 4;
                        for (int k = 0; k < K; ++k) dst[(size_t)(k >> 2) * t.nrb * 32 + (k & 3)] = vr[k];
                    }
                }
                int k0 = 0, nst = 0;
                do {  // a job with K == 0 still gets one (empty) stage that carries its record
                    int nk = std::max(0, std::min(kc, K4 - k0));
                    const size_t sb = srec.size();
                    srec.resize(sb + SV_SREC, 0);
                    int npieces = 0;
                    int xbytes = 0;
                    for (;;) {  // shrink the stage until its copies fit the record
                        npieces = xbytes = 0;
                        bool too_many = false;
                        auto add_piece = [&](int src, int slot, int code, int bytes) {
                            if (npieces >= SV_MAXRUNS) { too_many = true; return; }
                            srec[sb + 6 + 2 * npieces] = src;
                            srec[sb + 7 + 2 * npieces] = (slot << 8) | code;
                            ++npieces;
                            xbytes += bytes;
                        };
                        for (int pl = 0; pl < nsrc && !too_many; ++pl) {
                            const int32_t* ip = (pl == 0 ? p.i0 : (pl == 1 ? p.i1 : p.i2)) + p.blk_iptr[b] + t.k_lo;
                            auto row_of = [&](int k) { return (k0 + k < K) ? ip[k0 + k] : zrow; };
                            int run_src = -1, run_slot = 0, run_groups = 0;  // run of whole 4-row groups of consecutive rows
                            auto flush = [&]() {
                                for (int lg = 3; lg >= 0 && !too_many; --lg)  // boxes of 32, 16, 8, 4 rows
                                    while (run_groups >= (1 << lg) && !too_many) {
                                        add_piece(run_src, run_slot, lg + 2, (4 << lg) * (stride_w + 4) * 8);
                                        run_src += 4 << lg; run_slot += 4 << lg; run_groups -= 1 << lg;
                                    }
                                run_groups = 0;
                            };
                            for (int g = 0; g < nk && !too_many; g += 4) {
                                // absent / padded rows (index zrow) lie past the end of the tensor and read as zeros; a group
                                // of them is "consecutive" from any OOB row, a mixed group is not
                                int r0 = row_of(g);
                                bool consecutive = true, all_absent = (r0 == zrow);
                                for (int k = 1; k < 4; ++k) {
                                    const int rk = row_of(g + k);
                                    all_absent = all_absent && rk == zrow;
                                    consecutive = consecutive && rk == r0 + k && rk != zrow && r0 != zrow;
                                }
                                if (all_absent) { consecutive = true; r0 = (run_groups > 0 && run_src + 4 * run_groups >= zrow) ? run_src + 4 * run_groups : zrow; }
                                if (consecutive) {
                                    if (run_groups > 0 && r0 == run_src + 4 * run_groups) { ++run_groups; continue; }
                                    flush();
                                    run_src = r0; run_slot = pl * kc + g; run_groups = 1;
                                } else {
                                    flush();
                                    for (int k = 0; k < 4 && !too_many; ++k)
                                        add_piece(row_of(g + k), pl * kc + g + k, 128, stride_w * 8);
                                }
                            }
                            if (!too_many) flush();
                        }
                        if (!too_many) break;
                        if (nk <= 4) return fail(h, FCB_ERR_INVALID, "internal: a 4-row stage needs more than %d copies", SV_MAXRUNS);
                        nk -= 4;
                    }
                    srec[sb + 0] = nk;
                    srec[sb + 1] = npieces;
                    srec[sb + 2] = (int)((v0 + (size_t)k0 * 8 * t.nrb) / 32);
                    srec[sb + 3] = nk * 64 * t.nrb;
                    srec[sb + 4] = k0 == 0 ? (int)(jb / SV_JREC) : -1;
                    srec[sb + 5] = xbytes;
                    k0 += std::max(nk, 4);
                    ++nst;
                } while (k0 < K4);
                jrec[jb + 7] = nst;
            }
        }
        cta_sptr.push_back((int)(srec.size() / SV_SREC));  // one past the last CTA of this launch
        cta_jptr.push_back((int)(jrec.size() / SV_JREC));
        d.launches.push_back(L);
    }
    d.nstages_total = (long long)(srec.size() / SV_SREC);
    d.njobs = (int)(jrec.size() / SV_JREC);
    d.packed_doubles = (long long)packed.size();
    if (packed.size() / 32 > 0x7fffffffULL) return fail(h, FCB_ERR_INVALID, "factor too large for 32-bit tile offsets");
    if (srec.empty()) srec.resize(SV_SREC, 0);
    if (jrec.empty()) jrec.resize(SV_JREC, -1);
    if (packed.empty()) packed.resize(32, 0.0);
    if (cta_sptr.empty()) { cta_sptr.push_back(0); cta_jptr.push_back(0); }
    TRY(upload(h, &d.srec, srec.data(), srec.size()));
    TRY(upload(h, &d.jrec, jrec.data(), jrec.size()));
    TRY(upload(h, &d.cta_sptr, cta_sptr.data(), cta_sptr.size()));
    TRY(upload(h, &d.cta_jptr, cta_jptr.data(), cta_jptr.size()));
    TRY(upload(h, &d.vals, packed.data(), packed.size()));
    d.asm_n = p.asm_n;
    d.asm_lptr.assign((size_t)p.nlaunch + 1, 0);
    for (int l = 0; l <= p.nlaunch; ++l) {
        d.asm_lptr[l] = p.asm_lptr ? p.asm_lptr[l] : (l <= p.n_forward_launches ? 0 : p.asm_n);
        if (d.asm_lptr[l] < 0 || d.asm_lptr[l] > p.asm_n || (l > 0 && d.asm_lptr[l] < d.asm_lptr[l - 1]))
            return fail(h, FCB_ERR_INVALID, "plan: asm_lptr is malformed");
    }
    if (p.asm_n > 0) {
        for (int i = 0; i < p.asm_n; ++i) {
            if (p.asm_dst[i] < 0 || p.asm_dst[i] >= zrow || p.asm_ptr[i + 1] < p.asm_ptr[i])
                return fail(h, FCB_ERR_INVALID, "plan gather-sum row %d is malformed", i);
            for (int k = p.asm_ptr[i]; k < p.asm_ptr[i + 1]; ++k)
                if (p.asm_src[k] < 0 || p.asm_src[k] >= zrow) return fail(h, FCB_ERR_INVALID, "plan gather-sum source out of range");
        }
        TRY(upload(h, &d.asm_ptr, p.asm_ptr, (size_t)p.asm_n + 1));
        TRY(upload(h, &d.asm_src, p.asm_src, (size_t)std::max(1, p.asm_ptr[p.asm_n])));
        TRY(upload(h, &d.asm_dst, p.asm_dst, (size_t)p.asm_n));
    }
    CK(cudaStreamSynchronize(h->stream));  // host vectors go out of scope
    return FCB_OK;
}


// Compile the subtree clusters of a plan into device programs (see k_cluster_sweep) and upload them.
int upload_clusters(fcb_context* h, DevPlan& d, const fcb_plan& p, const int32_t* perm) {
    d.ncluster = p.ncluster;
    d.tiers.clear();
    if (p.ncluster <= 0) return FCB_OK;
    if (p.ntier <= 0 || !p.tier_ptr || p.tier_ptr[0] != 0 || p.tier_ptr[p.ntier] != p.ncluster)
        return fail(h, FCB_ERR_INVALID, "cluster plan: tier_ptr is malformed");
    const int zrow = 2 * p.n + p.nU;
    std::vector<int> prog;                 // all programs
    std::vector<int4> loc[2];
    std::vector<double> vals;              // A fragments, 32 doubles each
    int srow_max = 0, prog_max = 0;
    for (int q = 0; q < p.ncluster; ++q) {
        const int f0 = p.cl_fptr[q], f1 = p.cl_fptr[q + 1], nf = f1 - f0;
        if (nf <= 0 || f1 > p.nfront) return fail(h, FCB_ERR_INVALID, "cluster %d has no fronts", q);
        std::vector<int> off(nf + 1, 0);
        for (int k = 0; k < nf; ++k) {
            const int f = f0 + k;
            if (p.fr_w[f] < 0 || p.fr_m[f] < 0 || p.fr_c0[f] < 0 || p.fr_c0[f] + p.fr_w[f] > p.n || p.fr_sptr[f + 1] - p.fr_sptr[f] != p.fr_m[f] ||
                (k > 0 && p.fr_c0[f] < p.fr_c0[f - 1] + p.fr_w[f - 1]))
                return fail(h, FCB_ERR_INVALID, "cluster %d: front %d is malformed", q, f);
            off[k + 1] = off[k] + p.fr_w[f];
        }
        const int nown = off[nf];
        const int32_t* rs = p.fr_struct + p.fr_sptr[f1 - 1];
        const int mroot = p.fr_m[f1 - 1];
        const int srows = nown + mroot;
        if (srows > 65000) return fail(h, FCB_ERR_INVALID, "cluster %d: %d resident rows do not fit 16-bit tables", q, srows);
        srow_max = std::max(srow_max, srows);
        auto local = [&](int g) -> int {  // solver row -> resident row
            int lo = 0, hi = nf;  // last front with c0 <= g
            while (lo < hi) {
                const int mid = (lo + hi) / 2;
                if (p.fr_c0[f0 + mid] <= g) lo = mid + 1;
                else hi = mid;
            }
            if (lo > 0 && g < p.fr_c0[f0 + lo - 1] + p.fr_w[f0 + lo - 1]) return off[lo - 1] + g - p.fr_c0[f0 + lo - 1];
            const int32_t* it = std::lower_bound(rs, rs + mroot, g);
            if (it != rs + mroot && *it == g) return nown + (int)(it - rs);
            return -1;
        };
        for (int dir = 0; dir < 2; ++dir) {
            std::vector<int> ops;
            std::vector<unsigned short> tab;  // 16-bit tables: output rows of the forward operations
            std::vector<unsigned> ktab;       // byte offsets of the gathered rows, [k-step / 4][lane % 4][4] per operation
            const long long frag0 = (long long)(vals.size() / 32);
            // gathered rows of an operation -> ktab; returns its offset (a multiple of 16 entries)
            auto add_ktab = [&](const std::vector<int>& rows) {
                const int koff = (int)ktab.size(), ng = ((int)rows.size() + 15) / 16;
                ktab.resize(ktab.size() + (size_t)ng * 16, 0u);
                for (size_t kq = 0; kq < rows.size(); ++kq) {
                    const size_t ks = kq >> 2, tg = kq & 3;
                    ktab[koff + (ks >> 2) * 16 + tg * 4 + (ks & 3)] = (unsigned)rows[kq] * (unsigned)(CL_XS * 8);
                }
                return koff;
            };
            auto add_op = [&](int M, int K, int mode, int koff, int oarg, const double* V, int ldv, int r0, double sign) {
                const int nrb = (M + 7) / 8, nks = (K + 3) / 4, cks = std::max(4, (CL_CHUNK / nrb) & ~3);
                // few units and a long K: split K over the idle warps (parts take the groups of four k-steps round-robin)
                int kp = 1;
                while (kp < 4 && 2 * nrb * (kp * 2) <= CL_NW && (nks + 3) / 4 >= 2 * (kp * 2)) kp *= 2;
                if ((kp - 1) * 2 * nrb * 1024 > CL_RED_BYTES) kp = 1;
                const int rec[CL_OPREC] = {M, nks, nrb, mode, koff, oarg, kp, cks};
                ops.insert(ops.end(), rec, rec + CL_OPREC);
                const size_t v0 = vals.size();
                vals.resize(v0 + (size_t)nks * nrb * 32, 0.0);
                for (int r = 0; r < M; ++r) {
                    const double* vr = V + (size_t)(r0 + r) * ldv;
                    double* dst = vals.data() + v0 + (size_t)(r >> 3) * 32 + (size_t)(r & 7) * 4;
                    for (int k = 0; k < K; ++k) dst[(size_t)(k >> 2) * nrb * 32 + (k & 3)] = sign * vr[k];
                }
            };
            for (int kk = 0; kk < nf; ++kk) {
                const int k = dir == 0 ? kk : nf - 1 - kk, f = f0 + k;
                const int w = p.fr_w[f], m = p.fr_m[f];
                const int32_t* st = p.fr_struct + p.fr_sptr[f];
                if (w == 0) continue;
                if (dir == 0) {
                    if (m == 0) continue;
                    std::vector<int> krows(w);
                    for (int kq = 0; kq < w; ++kq) krows[kq] = off[k] + kq;
                    const int koff = add_ktab(krows);
                    for (int r0 = 0; r0 < m; r0 += CL_MAXROWS) {
                        const int M = std::min(CL_MAXROWS, m - r0);
                        const int ooff = (int)tab.size();
                        for (int r = 0; r < ((M + 7) & ~7); ++r) {
                            int lr = 0xffff;
                            if (r < M) {
                                lr = local(st[r0 + r]);
                                if (lr < 0) return fail(h, FCB_ERR_INVALID, "cluster %d: boundary row %d of front %d is not resident", q, st[r0 + r], f);
                            }
                            tab.push_back((unsigned short)lr);
                        }
                        add_op(M, w, 0, koff, ooff, p.cl_vals + p.fr_eptr[f], w, r0, -1.0);
                    }
                } else {
                    if (w > CL_MAXROWS) return fail(h, FCB_ERR_INVALID, "cluster %d: front %d is wider than %d", q, f, CL_MAXROWS);
                    const int K = w + m;
                    std::vector<int> krows(K);
                    for (int kq = 0; kq < K; ++kq) {
                        krows[kq] = kq < w ? off[k] + kq : local(st[kq - w]);
                        if (krows[kq] < 0) return fail(h, FCB_ERR_INVALID, "cluster %d: boundary row %d of front %d is not resident", q, st[kq - w], f);
                    }
                    add_op(w, K, 1, add_ktab(krows), off[k], p.cl_vals + p.fr_bptr[f], K, 0, 1.0);
                }
            }
            // imports (forward): sorted by consumer warp (resident row mod CL_NW), original order inside
            std::vector<std::array<int, 2>> imp;
            int wfirst[CL_NW + 1] = {0};
            if (dir == 0) {
                std::vector<std::array<int, 3>> tmp;
                for (long long t = p.cl_iptr[q]; t < p.cl_iptr[q + 1]; ++t) {
                    const int lr = local(p.imp_dst[t]);
                    if (lr < 0 || p.imp_src[t] < 0 || p.imp_src[t] >= zrow) return fail(h, FCB_ERR_INVALID, "cluster %d: import %lld is malformed", q, t);
                    tmp.push_back({lr % CL_NW, p.imp_src[t], lr});
                }
                std::stable_sort(tmp.begin(), tmp.end(), [](const std::array<int, 3>& x, const std::array<int, 3>& y) { return x[0] < y[0]; });
                for (auto& e : tmp) { imp.push_back({e[1], e[2]}); ++wfirst[e[0] + 1]; }
                for (int wq = 0; wq < CL_NW; ++wq) wfirst[wq + 1] += wfirst[wq];
            }
            // ---- lay the program out
            const size_t base = prog.size();
            std::vector<int> pr(CL_HDR, 0);
            pr[0] = (int)(ops.size() / CL_OPREC); pr[1] = nown; pr[2] = mroot; pr[3] = (int)imp.size();
            pr[4] = dir == 0 ? p.cl_ustore[q] : -1;
            if (dir == 0 && p.cl_ustore[q] >= 0 && (p.cl_ustore[q] < 2 * p.n || p.cl_ustore[q] + mroot > zrow))
                return fail(h, FCB_ERR_INVALID, "cluster %d: update vector outside the U region", q);
            if (frag0 > 0x7fffffffLL) return fail(h, FCB_ERR_INVALID, "cluster factor too large for 32-bit fragment offsets");
            pr[5] = (int)frag0;
            pr[11] = (dir == 1 && p.cl_iptr[q + 1] > p.cl_iptr[q]) ? 1 : 0;  // x rows go to Z only if a lower cluster gathers them
            for (int wq = 0; wq <= CL_NW; ++wq) pr[12 + wq] = wfirst[wq];
            pr[6] = (int)pr.size();  // grow
            for (int k = 0; k < nf; ++k)
                for (int r = 0; r < p.fr_w[f0 + k]; ++r) pr.push_back(p.fr_c0[f0 + k] + r);
            pr[30] = 1;  // the forward sweep stores y
            pr[29] = (int)pr.size();  // gdof (backward)
            if (dir == 1)
                for (int k = 0; k < nf; ++k)
                    for (int r = 0; r < p.fr_w[f0 + k]; ++r) pr.push_back(perm[p.fr_c0[f0 + k] + r]);
            while (pr.size() % 4) pr.push_back(0);
            pr[7] = (int)pr.size();  // ops (16-byte aligned: read as int4)
            const int kbase = pr[7] + (int)ops.size();  // gathered-row tables right behind (16-byte aligned: read as uint4)
            for (size_t o = 0; o < ops.size(); o += CL_OPREC) ops[o + 4] += kbase;
            pr.insert(pr.end(), ops.begin(), ops.end());
            pr.insert(pr.end(), ktab.begin(), ktab.end());
            pr[8] = (int)pr.size();  // 16-bit tables
            pr.resize(pr.size() + (tab.size() + 1) / 2, 0);
            memcpy(pr.data() + pr[8], tab.data(), tab.size() * sizeof(unsigned short));
            pr[9] = (int)pr.size();  // aux
            if (dir == 0) for (auto& e : imp) { pr.push_back(e[0]); pr.push_back(e[1]); }
            else for (int j = 0; j < mroot; ++j) {
                if (rs[j] < 0 || rs[j] >= p.n) return fail(h, FCB_ERR_INVALID, "cluster %d: boundary row out of range", q);
                pr.push_back(rs[j]);
            }
            while (pr.size() % 4) pr.push_back(0);
            pr[10] = (int)pr.size();
            prog_max = std::max(prog_max, (int)pr.size());
            bool one_run = true;
            for (int k = 1; k < nf; ++k) one_run = one_run && p.fr_c0[f0 + k] == p.fr_c0[f0 + k - 1] + p.fr_w[f0 + k - 1];
            loc[dir].push_back(make_int4((int)base, (int)pr.size(), one_run ? p.fr_c0[f0] : -1, nown));
            prog.insert(prog.end(), pr.begin(), pr.end());
        }
    }
    d.cl_srow_max = srow_max;
    d.cl_prog_max = prog_max;
    d.cl_smem = srow_max * CL_XS * 8 + CL_RING_BYTES + CL_RED_BYTES + prog_max * 4 + 8 * (1 + 2 * CL_STAGES);
    d.cl_frags = (long long)(vals.size() / 32);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, h->device));
    if (d.cl_smem > (int)prop.sharedMemPerBlockOptin)
        return fail(h, FCB_ERR_INVALID, "subtree clusters need %d bytes of shared memory per CTA (%d resident rows, %d-word programs), the device offers %d: "
                    "build the plan with a smaller cluster_rows", d.cl_smem, srow_max, prog_max, (int)prop.sharedMemPerBlockOptin);
    for (int t = 0; t < p.ntier; ++t) d.tiers.push_back({p.tier_ptr[t], p.tier_ptr[t + 1] - p.tier_ptr[t]});
    if (vals.empty()) vals.resize(32, 0.0);
    TRY(upload(h, &d.cprog, prog.data(), prog.size()));
    TRY(upload(h, &d.cprog_loc[0], loc[0].data(), loc[0].size()));
    TRY(upload(h, &d.cprog_loc[1], loc[1].data(), loc[1].size()));
    TRY(upload(h, &d.cvals, vals.data(), vals.size()));
    CK(cudaStreamSynchronize(h->stream));
    if (getenv("FCB_VERBOSE"))
        fprintf(stderr, "[fcb200] clusters: %d in %d tiers, max %d resident rows, programs <= %d words, %lld A fragments, %d B smem per CTA\n",
                p.ncluster, p.ntier, srow_max, prog_max, d.cl_frags, d.cl_smem);
    return FCB_OK;
}

void fill_tables(double phi[7][6], double dphi[7][6][2], double w[7], double mass[6][6]) {
    const double s15 = std::sqrt(15.0);
    const double a1 = (6.0 - s15) / 21.0, a2 = (6.0 + s15) / 21.0;
    const double w1 = (155.0 - s15) / 2400.0, w2 = (155.0 + s15) / 2400.0;
    const double xi[7] = {1.0 / 3.0, a1, 1 - 2 * a1, a1, a2, 1 - 2 * a2, a2};
    const double eta[7] = {1.0 / 3.0, a1, a1, 1 - 2 * a1, a2, a2, 1 - 2 * a2};
    const double ww[7] = {9.0 / 80.0, w1, w1, w1, w2, w2, w2};
    const double dl[3][2] = {{-1, -1}, {1, 0}, {0, 1}};
    const int ed[3][2] = {{1, 2}, {0, 2}, {0, 1}};
    for (int q = 0; q < 7; ++q) {
        const double l[3] = {1.0 - xi[q] - eta[q], xi[q], eta[q]};
        w[q] = ww[q];
        for (int i = 0; i < 3; ++i) {
            phi[q][i] = l[i] * (2 * l[i] - 1);
            for (int k = 0; k < 2; ++k) dphi[q][i][k] = (4 * l[i] - 1) * dl[i][k];
        }
        for (int e = 0; e < 3; ++e) {
            const int i = ed[e][0], j = ed[e][1];
            phi[q][3 + e] = 4 * l[i] * l[j];
            for (int k = 0; k < 2; ++k) dphi[q][3 + e][k] = 4 * (l[i] * dl[j][k] + l[j] * dl[i][k]);
        }
    }
    for (int a = 0; a < 6; ++a)
        for (int b = 0; b < 6; ++b) {
            double s = 0;
            for (int q = 0; q < 7; ++q) s += w[q] * phi[q][a] * phi[q][b];
            mass[a][b] = s;
        }
}

// ---- enqueue helpers ---------------------------------------------------------------------------
struct PhaseMark {
    fcb_context* h;
    bool on;
    void mark(int i) {
        if (on) cudaEventRecord(h->ev[i], h->stream);
    }
};

// Crank-Nicolson: the rhs rows first receive E u (k_spmm); the element pass then adds -N(u) (nsforms.py:219-229)
int enqueue_spmm(fcb_context* h, const double* u) {
    if (h->sp_csr) {
        dim3 grid((h->n + 7) / 8, h->ldb / 32), block(32, 8);
        k_spmm<<<grid, block, 0, h->stream>>>(h->n, h->cn_ptr, h->cn_idx, h->cn_val, u, h->Z, h->ldb);
    } else {
        // items = (block, slice) with the slice fastest: the CTAs working on one block run together and its panels come from HBM once
        SpmmArgs a{h->sp_ucols, h->sp_ginfo, h->sp_kslots, h->sp_avals, u, h->Z, h->ldb, h->ldb / 32};
        k_spmm_mma<<<h->sp_nblk * (h->ldb / 32), dim3(32, SPMM_WARPS), h->sp_smem, h->stream>>>(a);
    }
    h->launches += 1;
    CK(cudaGetLastError());
    return FCB_OK;
}

// element pass on the state u: b <- b(u) and either a <- a(u) (bprev == nullptr) or, fused, the next step's
// right-hand side rows Z[0,n) <- a(u) + bprev
int enqueue_element(fcb_context* h, const double* u, double* a, double* b, const double* bprev) {
    PatchArgs p;
    p.mode = h->scheme == 1 ? 2 : (bprev ? 1 : 0);
    p.pcell_ptr = h->pcell_ptr; p.pcnode = h->pcnode; p.pgeo = h->pgeo; p.plnode = h->plnode;
    p.pnode_ptr = h->pnode_ptr; p.pnode_dst = h->pnode_dst; p.psrc = h->psrc; p.pacc_rows = h->pacc_rows;
    p.u = u; p.a = a; p.b = b; p.scratch = h->pscratch; p.epart = h->epart;
    p.bprev = bprev; p.Zb = h->Z; p.prow = h->prow;
    p.nN = h->nN; p.nV = h->nV; p.ldb = h->ldb; p.tab_off = h->patch_tab_off;
    p.ca = 2.0 / h->dt; p.cb = -0.5 / h->dt; p.na = -2.0; p.nb = 1.0;  // BDF2: rhs = a_n + b_{n-1}
    if (h->scheme == 1) { p.ca = 0.0; p.cb = 0.0; p.na = -1.0; p.nb = 0.0; }
    dim3 grid(h->npatch, h->ldb / 32), block(32, EP_WARPS);
    if (h->nonlinear) k_element_patch<true><<<grid, block, h->patch_smem, h->stream>>>(p);
    else k_element_patch<false><<<grid, block, h->patch_smem, h->stream>>>(p);
    h->launches += 1;
    if (h->nshared > 0) {
        dim3 g2((h->nshared + 7) / 8, h->ldb / 32), b2(32, 8);
        k_patch_merge<<<g2, b2, 0, h->stream>>>(h->nshared, h->mptr, h->msrc, h->mnode, h->mrow, h->pscratch, a, b, bprev, h->Z,
                                                p.mode, h->nN, h->ldb);
        h->launches += 1;
    }
    CK(cudaGetLastError());
    return FCB_OK;
}

int enqueue_measure(fcb_context* h, const double* up, bool roll_ctrl = false) {
    dim3 grid(h->ldb / 32), block(32, MEAS_WARPS);
    k_measure<<<grid, block, 0, h->stream>>>(h->ns, h->sensor_ptr, h->sensor_idx, h->sensor_val, up, h->y, h->epart,
                                             h->nblk_total, h->dE, h->na, h->uctrl, roll_ctrl ? h->uctrl_prev : nullptr, h->ldb);
    h->launches += 1;
    CK(cudaGetLastError());
    return FCB_OK;
}


constexpr size_t DBG_PER_LAUNCH = 8 * 8192;  // 8 timeline slots for up to 8192 CTAs

template <int NWC, bool KS>
void launch_sweep(fcb_context* h, const DevPlan& pl, const DevPlan::Launch& L, int mapset, double* xout, unsigned long long* dbg) {
    // programmatic dependent launch: the CTAs of level l+1 may start (barrier init, stream-record prefetch) while
    // level l drains; they wait on griddepcontrol before touching Z
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(L.grid, L.nslab);
    cfg.blockDim = dim3(32, KS ? 5 : NWC + 1);
    cfg.dynamicSmemBytes = SweepCfg<NWC>::smem_bytes(L.nstages, L.slots) + (KS ? SV_RED_BYTES : 0);
    cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = h->use_pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, k_front_sweep<NWC, KS>, h->zmaps[mapset], (const int*)pl.srec, (const int*)pl.jrec,
                       (const int*)(pl.cta_sptr + L.cta_off), (const int*)(pl.cta_jptr + L.cta_off), (const double*)pl.vals, h->Z,
                       h->ldb, L.nstages, L.slots, xout, h->diverged, h->Nv, dbg);
}

void launch_clusters(fcb_context* h, const DevPlan& pl, const DevPlan::Tier& t, int backward, double* xout) {
    ClusterArgs a;
    const int tier_index = (int)(&t - pl.tiers.data());
    a.dbg = (h->cl_dbg && tier_index < 8) ? h->cl_dbg + (size_t)(tier_index * 2 + backward) * DBG_PER_LAUNCH : nullptr;
    a.prog = pl.cprog; a.prog_loc = pl.cprog_loc[backward] + t.first; a.vals = pl.cvals; a.Z = h->Z; a.xout = xout;
    a.diverged = h->diverged; a.n = pl.n; a.Nv = h->Nv; a.ldb = h->ldb; a.nslab = h->ldb / CL_W; a.backward = backward;
    a.prog_max_words = pl.cl_prog_max; a.srow_max = pl.cl_srow_max;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(t.count * a.nslab);
    cfg.blockDim = dim3(32, CL_NW + 1);
    cfg.dynamicSmemBytes = pl.cl_smem;
    cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = h->use_pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, k_cluster_sweep, a);
    h->launches += 1;
}

int enqueue_solve(fcb_context* h, const DevPlan& pl, double* xout, PhaseMark* pm) {
    for (const DevPlan::Tier& t : pl.tiers) launch_clusters(h, pl, t, 0, xout);  // forward: the subtree clusters first, tier by tier
    for (int l = 0; l < pl.nlaunch; ++l) {
        if (pm && l == pl.n_forward) pm->mark(FCB_PHASE_BACKWARD);
        if (const int nsum = pl.asm_lptr[l + 1] - pl.asm_lptr[l]; nsum > 0) {
            // gather-sums of this launch: virtual update vectors of fronts with more than two children, top right-hand side
            const int r0 = pl.asm_lptr[l];
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3((nsum + 7) / 8, h->ldb / 32);
            cfg.blockDim = dim3(32, 8);
            cfg.stream = h->stream;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr;
            cfg.numAttrs = h->use_pdl ? 1 : 0;
            cudaLaunchKernelEx(&cfg, k_gather_sum, nsum, (const int*)(pl.asm_ptr + r0), (const int*)pl.asm_src, (const int*)(pl.asm_dst + r0), h->Z, h->ldb);
            h->launches += 1;
        }
        const DevPlan::Launch& L = pl.launches[l];
        if (L.grid <= 0) continue;
        unsigned long long* dbg = (pm && h->sweep_dbg) ? h->sweep_dbg + (size_t)l * DBG_PER_LAUNCH : nullptr;
        switch (L.nwc) {
            case 8: launch_sweep<8, false>(h, pl, L, 3, xout, dbg); break;
            case 4: launch_sweep<4, false>(h, pl, L, 2, xout, dbg); break;
            case 2: launch_sweep<2, false>(h, pl, L, 1, xout, dbg); break;
            default:
                if (L.ksplit) launch_sweep<1, true>(h, pl, L, 0, xout, dbg);
                else launch_sweep<1, false>(h, pl, L, 0, xout, dbg);
                break;
        }
        h->launches += 1;
    }
    if (pm && pl.n_forward >= pl.nlaunch) pm->mark(FCB_PHASE_BACKWARD);
    for (size_t t = pl.tiers.size(); t-- > 0;) launch_clusters(h, pl, pl.tiers[t], 1, xout);  // backward: clusters last, top tier first
    CK(cudaGetLastError());
    return FCB_OK;
}

// one step with the current (order, parity); u_ctrl already in h->uctrl
// One step.  rhs_ready: Z[0,n) already holds a_n + b_{n-1} from the previous step's fused element pass and only the
// control terms are missing; otherwise (first step after fcb_set_state) the right-hand side is built from a and b.
int enqueue_step(fcb_context* h, int order, int parity, bool rhs_ready, PhaseMark* pm) {
    const DevPlan& pl = h->plan[order - 1];
    double* nxt = h->up[1 - parity];
    if (pm) pm->mark(FCB_PHASE_RHS);
    if (rhs_ready && order == 2) {
        if (h->ncrow > 0 && h->na > 0) {
            dim3 grid((h->ncrow + 7) / 8, h->ldb / 32), block(32, 8);
            k_ctrl_add<<<grid, block, 0, h->stream>>>(h->ncrow, h->crow, h->ccoef, h->na, h->uctrl, h->Z, h->ldb);
            h->launches += 1;
        }
        if (h->scheme == 1 && h->ncrow_prev > 0 && h->na > 0) {
            // Crank-Nicolson averages the body force over the step: + F u_ctrl^n / 2 (nsforms.py:226-229)
            dim3 grid((h->ncrow_prev + 7) / 8, h->ldb / 32), block(32, 8);
            k_ctrl_add<<<grid, block, 0, h->stream>>>(h->ncrow_prev, h->crow_prev, h->ccoef_prev, h->na, h->uctrl_prev, h->Z, h->ldb);
            h->launches += 1;
        }
    } else {
        dim3 grid((h->n + 7) / 8, h->ldb / 32), block(32, 8);
        k_rhs_build<<<grid, block, 0, h->stream>>>(h->n, h->Nv, h->perm, h->avec, h->bvec[1 - parity], order, h->na,
                                                   h->ctrl_rhs[order - 1], h->uctrl, h->Z, h->ldb);
        h->launches += 1;
    }
    if (pm) pm->mark(FCB_PHASE_FORWARD);
    TRY(enqueue_solve(h, pl, nxt, pm));
    if (pm) pm->mark(FCB_PHASE_POST);
    if (h->nbc > 0) {
        dim3 grid((h->nbc + 7) / 8, h->ldb / 32), block(32, 8);
        k_bc_fill<<<grid, block, 0, h->stream>>>(h->nbc, h->bc_dofs, h->na, h->bc_shape, h->uctrl, nxt, h->ldb);
        h->launches += 1;
    }
    if (pm) pm->mark(FCB_PHASE_SPMM);
    if (h->scheme == 1) TRY(enqueue_spmm(h, nxt));
    if (pm) pm->mark(FCB_PHASE_ELEMENT);
    // b(u_new) replaces b_{n-1}; the rhs of the next step = a(u_new) + b_n, with b_n = bvec[parity]
    TRY(enqueue_element(h, nxt, h->avec, h->bvec[1 - parity], h->bvec[parity]));
    if (pm) pm->mark(FCB_PHASE_MEASURE);
    TRY(enqueue_measure(h, nxt, h->scheme == 1 && h->na > 0));
    if (pm) pm->mark(FCB_NPHASES);
    CK(cudaGetLastError());
    return FCB_OK;
}

int enqueue_controller(fcb_context* h, int xparity) {
    k_controller<<<h->ldb / 32, dim3(32, CTRL_ROWS), 0, h->stream>>>(
        h->nx, h->ny, h->nu, h->ns, h->na, h->Ad, h->Bd, h->Cd, h->Dd, h->Ky, h->Fu, h->y, h->xk[xparity],
        h->xk[1 - xparity], h->uctrl, h->ldb);
    h->launches += 1;
    CK(cudaGetLastError());
    return FCB_OK;
}

int enqueue_log(fcb_context* h) {
    k_log<<<1, 256, 0, h->stream>>>(h->na, h->ns, h->dE, h->uctrl, h->y, h->ctl, h->costs, h->ldb);
    h->launches += 1;
    CK(cudaGetLastError());
    return FCB_OK;
}

int enqueue_load_ctrl(fcb_context* h) {
    if (h->na <= 0) return FCB_OK;
    k_load_ctrl<<<h->ldb / 32, 32, 0, h->stream>>>(h->na, h->ctl, h->uctrl, h->ldb);
    h->launches += 1;
    CK(cudaGetLastError());
    return FCB_OK;
}

enum RunMode { RUN_STEP = 0, RUN_CLOSED = 1, RUN_OPEN = 2 };

int capture(fcb_context* h, cudaGraphExec_t* exec, int* nodes, int mode, int parity, int xparity) {
    const long long before = h->launches;
    CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    int rc = FCB_OK;
    if (mode == RUN_CLOSED) rc = enqueue_controller(h, xparity);
    if (mode == RUN_OPEN) rc = enqueue_load_ctrl(h);
    if (rc == FCB_OK) rc = enqueue_step(h, 2, parity, true, nullptr);
    if (rc == FCB_OK && mode != RUN_STEP) rc = enqueue_log(h);
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
    *nodes = (int)(h->launches - before);
    h->launches = before;  // capturing does not launch
    if (rc != FCB_OK) return rc;
    if (e != cudaSuccess) return fail(h, FCB_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
    CK(cudaGraphInstantiate(exec, graph, 0));
    CK(cudaGraphDestroy(graph));
    return FCB_OK;
}

// copy [rows, B] user array (host or device) into [rows, ldb] device array (or back)
int copy_in(fcb_context* h, double* dst, const double* src, int rows) {
    CK(cudaMemcpy2DAsync(dst, (size_t)h->ldb * sizeof(double), src, (size_t)h->B * sizeof(double),
                         (size_t)h->B * sizeof(double), rows, cudaMemcpyDefault, h->stream));
    return FCB_OK;
}
template <typename T>
int copy_out(fcb_context* h, T* dst, const T* src, int rows) {
    CK(cudaMemcpy2DAsync(dst, (size_t)h->B * sizeof(T), src, (size_t)h->ldb * sizeof(T), (size_t)h->B * sizeof(T), rows,
                         cudaMemcpyDefault, h->stream));
    return FCB_OK;
}

int run_one_step(fcb_context* h, int mode) {
    if (h->order == 2 && h->rhs_ready) {
        cudaGraphExec_t* exec = mode == RUN_CLOSED ? &h->g_loop[h->parity][h->xparity] : (mode == RUN_OPEN ? &h->g_open[h->parity] : &h->g_step[h->parity]);
        int* nodes = mode == RUN_CLOSED ? &h->g_loop_nodes[h->parity][h->xparity] : (mode == RUN_OPEN ? &h->g_open_nodes[h->parity] : &h->g_step_nodes[h->parity]);
        if (!*exec) TRY(capture(h, exec, nodes, mode, h->parity, h->xparity));
        CK(cudaGraphLaunch(*exec, h->stream));
        h->launches += *nodes;
    } else {
        if (mode == RUN_CLOSED) TRY(enqueue_controller(h, h->xparity));
        if (mode == RUN_OPEN) TRY(enqueue_load_ctrl(h));
        TRY(enqueue_step(h, h->order, h->parity, h->rhs_ready, nullptr));
        if (mode != RUN_STEP) TRY(enqueue_log(h));
    }
    h->parity ^= 1;
    if (mode == RUN_CLOSED) h->xparity ^= 1;
    h->order = 2;
    h->rhs_ready = true;
    return FCB_OK;
}

void destroy(fcb_context* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->pin_in) cudaFreeHost(h->pin_in);
    if (h->pin_out) cudaFreeHost(h->pin_out);
    for (int i = 0; i < 2; ++i) {
        if (h->g_step[i]) cudaGraphExecDestroy(h->g_step[i]);
        if (h->g_open[i]) cudaGraphExecDestroy(h->g_open[i]);
        for (int j = 0; j < 2; ++j)
            if (h->g_loop[i][j]) cudaGraphExecDestroy(h->g_loop[i][j]);
    }
    void* ptrs[] = {h->costs, h->sp_ucols, h->sp_ginfo, h->sp_kslots, h->sp_avals, h->cn_ptr, h->cn_idx, h->cn_val, h->uctrl_prev, h->ccoef_prev, h->crow_prev, h->crow, h->ccoef, h->bc_dofs, h->cell_nodes, h->perm, h->iperm, h->Jinv, h->detJ, h->bc_shape, h->ctrl_rhs[0],
                    h->ctrl_rhs[1], h->sensor_ptr, h->sensor_idx, h->sensor_val, h->up[0], h->up[1], h->avec,
                    h->bvec[0], h->bvec[1], h->Z, h->epart, h->uctrl, h->y, h->diverged, h->Ad, h->Bd, h->Cd,
                    h->Dd, h->Ky, h->Fu, h->xk[0], h->xk[1], h->series, h->useries, h->ctl, h->sweep_dbg, h->cl_dbg,
                    h->pcell_ptr, h->pcnode, h->pgeo, h->pnode_ptr, h->pnode_dst, h->mptr, h->msrc, h->mnode, h->plnode, h->pscratch, h->prow, h->mrow, h->psrc, h->pacc_rows};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    for (int i = 0; i < 2; ++i) {
        void* pp[] = {h->plan[i].srec, h->plan[i].jrec, h->plan[i].cta_sptr, h->plan[i].cta_jptr, h->plan[i].vals,
                      h->plan[i].asm_ptr, h->plan[i].asm_src, h->plan[i].asm_dst, h->plan[i].cprog, h->plan[i].cprog_loc[0],
                      h->plan[i].cprog_loc[1], h->plan[i].cvals};
        for (void* q : pp)
            if (q) cudaFree(q);
    }
    for (auto& e : h->ev)
        if (e) cudaEventDestroy(e);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

// Re-pack the Crank-Nicolson operator (CSR, rows in solver order, columns = canonical velocity dofs) for k_spmm_mma.
namespace spmm_pack {
typedef std::array<uint64_t, SPMM_CMAX / 64> Bits;
struct Group {
    int nm = 0, members[8];
    std::vector<int> slots;  // 4 per k-step, -1 = padding
};
struct Block {
    int r1 = 0;              // one past the last solver row covered
    std::vector<int> cols, rows;
    std::vector<Group> groups;
    int nk = 0, conflicts = 0;
};

// one block from the given (non-empty) rows; false if they touch more than SPMM_CMAX distinct columns
inline bool pack_block(const int32_t* ptr, const int32_t* idx, const std::vector<int>& rows, std::vector<int>& slot_of, Block& B) {
    B = Block();
    B.rows = rows;
    bool fits = true;
    for (int r : rows)
        for (int j = ptr[r]; j < ptr[r + 1] && fits; ++j)
            if (slot_of[idx[j]] < 0) {
                if ((int)B.cols.size() == SPMM_CMAX) { fits = false; break; }
                slot_of[idx[j]] = (int)B.cols.size();
                B.cols.push_back(idx[j]);
            }
    if (!fits) {
        for (int c : B.cols) slot_of[c] = -1;
        return false;
    }
    const int R = (int)B.rows.size();
    std::vector<Bits> bits((size_t)R);
    std::vector<int> order((size_t)R);
    for (int i = 0; i < R; ++i) {
        order[i] = i;
        bits[i].fill(0);
        for (int j = ptr[B.rows[i]]; j < ptr[B.rows[i] + 1]; ++j) { const int sl = slot_of[idx[j]]; bits[i][sl >> 6] |= 1ull << (sl & 63); }
    }
    for (int c : B.cols) slot_of[c] = -1;
    // seeds: the longest rows first (a vertex node and the edge nodes inside its neighbourhood make a tight panel);
    // members: the rows that add the fewest new columns, longer rows first among equals
    auto len = [&](int i) { return ptr[B.rows[i] + 1] - ptr[B.rows[i]]; };
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return len(x) > len(y); });
    std::vector<char> used((size_t)R, 0);
    for (int seed : order) {
        if (used[seed]) continue;
        Group G;
        Bits uni = bits[seed];
        G.members[G.nm++] = seed; used[seed] = 1;
        while (G.nm < 8) {
            int best = -1, best_cost = 1 << 30, best_len = -1;
            for (int i = 0; i < R; ++i) {
                if (used[i]) continue;
                int cost = 0;
                for (size_t q = 0; q < uni.size(); ++q) cost += __builtin_popcountll(bits[i][q] & ~uni[q]);
                if (cost < best_cost || (cost == best_cost && len(i) > best_len)) { best_cost = cost; best = i; best_len = len(i); }
            }
            if (best < 0) break;
            G.members[G.nm++] = best; used[best] = 1;
            for (size_t q = 0; q < uni.size(); ++q) uni[q] |= bits[best][q];
        }
        // k-steps: four columns each, one from each residue class of the slot (mod 4) while they last
        std::vector<int> bucket[4];
        for (int sl = 0; sl < (int)B.cols.size(); ++sl)
            if (uni[sl >> 6] >> (sl & 63) & 1) bucket[sl & 3].push_back(sl);
        size_t K = 0, take[4] = {0, 0, 0, 0};
        for (auto& bq : bucket) K += bq.size();
        const int nk = (int)((K + 3) / 4);
        for (int t = 0; t < nk; ++t) {
            int got = 0;
            bool clean = true;
            for (int q = 0; q < 4; ++q)
                if (take[q] < bucket[q].size()) { G.slots.push_back(bucket[q][take[q]++]); ++got; }
            while (got < 4) {
                size_t remaining = 0, left = 0;
                int q = -1;
                for (int qq = 0; qq < 4; ++qq) {
                    remaining += bucket[qq].size() - take[qq];
                    if (bucket[qq].size() - take[qq] > left) { left = bucket[qq].size() - take[qq]; q = qq; }
                }
                if (q < 0 || remaining <= (size_t)(nk - 1 - t) * 4) { G.slots.push_back(-1); ++got; continue; }  // fits later: pad
                G.slots.push_back(bucket[q][take[q]++]);  // a class ran dry: two columns of one class cost a bank conflict
                ++got;
                clean = false;
            }
            B.conflicts += !clean;
        }
        B.nk += nk;
        B.groups.push_back(std::move(G));
    }
    return true;
}

// Recursive coordinate bisection of the rows (by the position of their mesh node) into k compact clusters of equal
// size; the two velocity rows of a node stay together.  Compact 2-D clusters touch ~2.2 columns per row, consecutive
// solver rows (which follow separators, i.e. curves) ~3.1.
inline void bisect(std::vector<int>& rows, size_t lo, size_t hi, int k, const std::vector<int>& node, const double* xy,
                   std::vector<std::vector<int>>& out) {
    if (k <= 1 || hi - lo <= 1) {
        out.emplace_back(rows.begin() + lo, rows.begin() + hi);
        return;
    }
    double mn[2] = {1e300, 1e300}, mx[2] = {-1e300, -1e300};
    for (size_t i = lo; i < hi; ++i)
        for (int a = 0; a < 2; ++a) {
            const double v = xy[2 * node[rows[i]] + a];
            mn[a] = std::min(mn[a], v);
            mx[a] = std::max(mx[a], v);
        }
    const int ax = (mx[0] - mn[0] >= mx[1] - mn[1]) ? 0 : 1;
    std::sort(rows.begin() + lo, rows.begin() + hi, [&](int x, int y) {
        const double vx = xy[2 * node[x] + ax], vy = xy[2 * node[y] + ax];
        if (vx != vy) return vx < vy;
        if (node[x] != node[y]) return node[x] < node[y];
        return x < y;
    });
    const int k1 = k / 2;
    size_t cut = lo + (size_t)std::llround((double)(hi - lo) * k1 / k);
    while (cut > lo && cut < hi && node[rows[cut]] == node[rows[cut - 1]]) ++cut;
    cut = std::min(std::max(cut, lo + 1), hi - 1);
    bisect(rows, lo, cut, k1, node, xy, out);
    bisect(rows, cut, hi, k - k1, node, xy, out);
}
}  // namespace spmm_pack

int build_spmm(fcb_context* h, const fcb_problem* p) {
    using namespace spmm_pack;
    const int n = p->n_free;
    const int32_t *ptr = p->cn_ptr, *idx = p->cn_idx;
    const double* val = p->cn_val;
    for (int i = 0; i < n; ++i) {
        if (ptr[i + 1] - ptr[i] > SPMM_CMAX) return fail(h, FCB_ERR_INVALID, "Crank-Nicolson operator row %d has more than %d entries", i, SPMM_CMAX);
        for (int j = ptr[i] + 1; j < ptr[i + 1]; ++j)
            if (idx[j] <= idx[j - 1]) return fail(h, FCB_ERR_INVALID, "cn_idx must be strictly increasing within a row");
    }
    std::vector<int> ucols, ginfo;
    int nblk = 0;
    std::vector<unsigned short> kslots;
    std::vector<double> avals;
    std::vector<int> slot_of((size_t)h->Nv, -1);
    size_t nnz_total = 0, conflict_steps = 0, ngroups = 0, staged = 0, ksteps_total = 0;
    Block B;
    std::vector<int> live, node((size_t)n, 0);
    for (int r = 0; r < n; ++r) {
        if (ptr[r + 1] > ptr[r]) live.push_back(r);
        node[r] = p->perm[r] % h->nN;  // canonical dof -> mesh node (ux: node, uy: nN + node)
    }
    // without coordinates, the position of a node in the solver's nested-dissection order stands in for them
    std::vector<double> order_xy;
    const double* xy = p->node_xy;
    if (!xy) {
        order_xy.assign((size_t)2 * h->nN, 0.0);
        for (int r = n - 1; r >= 0; --r)
            if (ptr[r + 1] > ptr[r]) order_xy[2 * (size_t)node[r]] = (double)r;
        xy = order_xy.data();
    }
    std::vector<std::vector<int>> work, clusters;
    if (!live.empty()) bisect(live, 0, live.size(), (int)((live.size() + 8 * SPMM_WARPS - 3) / (8 * SPMM_WARPS - 2)), node, xy, work);
    while (!work.empty()) {  // halve the clusters that are too tall or touch too many columns
        std::vector<int> c = std::move(work.back());
        work.pop_back();
        if ((int)c.size() <= 8 * SPMM_WARPS && pack_block(ptr, idx, c, slot_of, B)) { clusters.push_back(std::move(c)); continue; }
        if (c.size() <= 1) return fail(h, FCB_ERR_INVALID, "Crank-Nicolson operator row %d does not fit a panel block", c.empty() ? -1 : c[0]);
        std::vector<std::vector<int>> halves;
        bisect(c, 0, c.size(), 2, node, xy, halves);
        for (auto& hc : halves) work.push_back(std::move(hc));
    }
    std::reverse(clusters.begin(), clusters.end());  // back to the spatial order of the bisection
    for (const std::vector<int>& c : clusters) {
        pack_block(ptr, idx, c, slot_of, B);
        ++nblk;
        ucols.insert(ucols.end(), B.cols.begin(), B.cols.end());
        ucols.resize((size_t)nblk * SPMM_CMAX, -1);  // fixed strides: no descriptor load in front of the gathers
        for (const Group& G : B.groups) {
            const int gnk = (int)G.slots.size() / 4;
            const size_t kbase = avals.size() / 32;  // first k-step of the group, a multiple of four
            ginfo.push_back((int)kbase);
            ginfo.push_back(gnk);
            for (int m = 0; m < 8; ++m) ginfo.push_back(m < G.nm ? B.rows[G.members[m]] : -1);
            ginfo.push_back(0);
            ginfo.push_back(0);
            const int gpad = (gnk + 3) & ~3;
            avals.resize((kbase + gpad) * 32, 0.0);
            kslots.resize((kbase + gpad) * 4, 0);
            ksteps_total += gnk;
            for (int t = 0; t < gnk; ++t) {
                const size_t base = (kbase + t) * 32;
                for (int c = 0; c < 4; ++c) {
                    const int sl = G.slots[t * 4 + c];
                    // padding columns: a zero panel column on some staged slot
                    kslots[((kbase + t) / 4 * 4 + c) * 4 + (t & 3)] = (unsigned short)(sl < 0 ? std::min(c, (int)B.cols.size() - 1) : sl);
                    if (sl < 0) continue;
                    const int col = B.cols[sl];
                    for (int m = 0; m < G.nm; ++m) {
                        const int row = B.rows[G.members[m]];
                        const int32_t* f = std::lower_bound(idx + ptr[row], idx + ptr[row + 1], col);
                        if (f != idx + ptr[row + 1] && *f == col) avals[base + m * 4 + c] = val[f - idx];
                    }
                }
            }
            for (int m = 0; m < G.nm; ++m) nnz_total += ptr[B.rows[G.members[m]] + 1] - ptr[B.rows[G.members[m]]];
        }
        for (size_t g = B.groups.size(); g < (size_t)SPMM_WARPS; ++g) {  // absent groups: no k-steps, no rows
            ginfo.push_back(0);
            ginfo.push_back(0);
            for (int m = 0; m < 10; ++m) ginfo.push_back(-1);
        }
        conflict_steps += B.conflicts;
        ngroups += B.groups.size();
        staged += B.cols.size();
    }
    // the kernel loads the first two chunks of a group unconditionally: two chunks of padding at the end
    avals.resize(avals.size() + 32 * 8, 0.0);
    kslots.resize(kslots.size() + 4 * 8, 0);
    h->sp_nblk = nblk;
    h->sp_smem = SPMM_SMEM;
    TRY(upload(h, &h->sp_ucols, ucols.data(), std::max<size_t>(ucols.size(), 1)));
    TRY(upload(h, &h->sp_ginfo, ginfo.data(), std::max<size_t>(ginfo.size(), 1)));
    TRY(upload(h, &h->sp_kslots, kslots.data(), std::max<size_t>(kslots.size(), 1)));
    TRY(upload(h, &h->sp_avals, avals.data(), std::max<size_t>(avals.size(), 1)));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaFuncSetAttribute(k_spmm_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, h->sp_smem));
    if (getenv("FCB_VERBOSE"))
        fprintf(stderr, "[fcb] spmm: %d blocks, %zu groups, %zu k-steps (%zu with a bank conflict), nnz %zu, panel fill %.2f, staged columns %zu (%.2f x Nv)\n",
                h->sp_nblk, ngroups, ksteps_total, conflict_steps, nnz_total,
                nnz_total ? (double)(ksteps_total * 32) / (double)nnz_total : 0.0, staged, (double)staged / h->Nv);
    return FCB_OK;
}

// Group the cells into patches for k_element_patch.  The nested-dissection numbering of the solver is a
// space-filling order of the mesh, so sorting cells by the lowest solver row among their nodes and cutting
// the list into chunks gives compact patches without needing coordinates.
int build_patches(fcb_context* h, const fcb_problem* p, const std::vector<int>& iperm) {
    const int nT = p->nT, nN = p->nN;
    int pc = 24;  // cells per patch: 4 sub-patches of 6 cells, ~94 accumulator rows x 1 KB: two CTAs per SM (measured 18..30)
    const char* env = getenv("FCB_PATCH_CELLS");
    if (env && atoi(env) >= EP_WARPS && atoi(env) <= 40) pc = atoi(env);
    // Cell order: recursive coordinate bisection of the cell centroids (longer extent, median split) when node
    // coordinates are given, so that consecutive runs of cells are compact blobs at every scale; without
    // coordinates, the solver's nested-dissection numbering (a space-filling order of the mesh) is used instead.
    std::vector<std::pair<int, int>> key(nT), patch_range;  // patch_range: [lo, hi) in the cell order
    if (p->node_xy) {
        std::vector<double> cx(nT), cy(nT);
        for (int e = 0; e < nT; ++e) {
            double sx = 0, sy = 0;
            for (int i = 0; i < 3; ++i) { sx += p->node_xy[2 * p->cell_nodes[e * 6 + i]]; sy += p->node_xy[2 * p->cell_nodes[e * 6 + i] + 1]; }
            cx[e] = sx / 3; cy[e] = sy / 3;
        }
        std::vector<int> order(nT);
        for (int e = 0; e < nT; ++e) order[e] = e;
        // k-way bisection with proportional cuts: ceil(nT / pc) patches of equal size (not a power of two of them),
        // then two median levels inside a patch order its cells into the four sub-patches
        struct Range { int lo, hi, k, below; };  // k: patches this range still splits into; below: levels below the patch level
        const int npatch_target = (nT + pc - 1) / pc;
        std::vector<Range> stack{{0, nT, npatch_target, npatch_target <= 1 ? 0 : -1}};
        if (npatch_target <= 1) patch_range.push_back({0, nT});
        while (!stack.empty()) {
            const Range r = stack.back();
            stack.pop_back();
            if (r.hi - r.lo <= 1 || r.below >= 2) continue;
            double x0 = 1e300, x1 = -1e300, y0 = 1e300, y1 = -1e300;
            for (int k = r.lo; k < r.hi; ++k) {
                x0 = std::min(x0, cx[order[k]]); x1 = std::max(x1, cx[order[k]]);
                y0 = std::min(y0, cy[order[k]]); y1 = std::max(y1, cy[order[k]]);
            }
            const std::vector<double>& c = (x1 - x0 >= y1 - y0) ? cx : cy;
            const int k1 = r.below < 0 ? r.k / 2 : 0;
            const int mid = r.below < 0 ? r.lo + (int)std::llround((double)(r.hi - r.lo) * k1 / r.k) : r.lo + (r.hi - r.lo) / 2;
            std::nth_element(order.begin() + r.lo, order.begin() + mid, order.begin() + r.hi,
                             [&](int a, int b) { return c[a] < c[b] || (c[a] == c[b] && a < b); });
            const int ks[2] = {k1, r.k - k1};
            const std::pair<int, int> halves[2] = {{r.lo, mid}, {mid, r.hi}};
            for (int hh = 0; hh < 2; ++hh) {
                int below = r.below >= 0 ? r.below + 1 : -1;
                if (r.below < 0 && ks[hh] <= 1) { below = 0; patch_range.push_back(halves[hh]); }
                stack.push_back({halves[hh].first, halves[hh].second, ks[hh], below});
            }
        }
        std::sort(patch_range.begin(), patch_range.end());
        for (int k = 0; k < nT; ++k) key[k] = {k, order[k]};
    } else {
        for (int e = 0; e < nT; ++e) {
            int k = INT_MAX;
            for (int i = 0; i < 6; ++i) {
                const int nd = p->cell_nodes[e * 6 + i];
                for (int c = 0; c < 2; ++c) {
                    const int r = iperm[nd + c * nN];
                    if (r >= 0) k = std::min(k, r);
                }
            }
            key[e] = {k, e};
        }
        std::sort(key.begin(), key.end());
        for (int lo = 0; lo < nT; lo += pc) patch_range.push_back({lo, std::min(nT, lo + pc)});
    }
    const int npatch = (int)patch_range.size();
    std::vector<int> pcell_ptr(1, 0), pcells, pcnode, pnode_ptr(1, 0), pnode_dst, node_npatch(nN, 0), pacc_rows;
    std::vector<double> pgeo;
    std::vector<unsigned char> plnode, psrc;
    std::vector<std::vector<int>> patch_nodes(npatch);
    std::vector<int> local(nN, -1), uniq(nN, -1);
    int max_rows = 0;
    for (int q = 0; q < npatch; ++q) {
        const int e0 = patch_range[q].first, e1 = patch_range[q].second;
        const int em = e0 + (e1 - e0) / 2;
        const int cut[EP_WARPS + 1] = {e0, e0 + (em - e0) / 2, em, em + (e1 - em) / 2, e1};  // the quarters of the bisection
        std::vector<int>& nodes = patch_nodes[q];
        std::vector<std::array<unsigned char, 4>> src;  // per unique node: its accumulator rows
        int rows = 0;
        for (int g = 0; g < EP_WARPS; ++g) {
            // sub-patch of warp g: a contiguous run of the (space-filling) cell order, with its own accumulator rows
            std::vector<int> touched;
            for (int k = cut[g]; k < cut[g + 1]; ++k) {
                const int e = key[k].second;
                pcells.push_back(e);
                for (int i = 0; i < 4; ++i) pgeo.push_back(p->Jinv[(size_t)e * 4 + i]);
                pgeo.push_back(p->detJ[e]);
                for (int i = 0; i < 6; ++i) {
                    const int nd = p->cell_nodes[e * 6 + i];
                    pcnode.push_back(nd);
                    if (local[nd] < 0) {
                        local[nd] = rows++;
                        touched.push_back(nd);
                        if (uniq[nd] < 0) { uniq[nd] = (int)nodes.size(); nodes.push_back(nd); src.push_back({255, 255, 255, 255}); }
                        auto& sl = src[uniq[nd]];
                        int kk = 0;
                        while (sl[kk] != 255) ++kk;  // at most one row per warp: kk < EP_WARPS
                        sl[kk] = (unsigned char)local[nd];
                    }
                    plnode.push_back((unsigned char)local[nd]);
                }
            }
            pcell_ptr.push_back((int)pcells.size());
            for (int nd : touched) local[nd] = -1;
        }
        if (rows > 254) return fail(h, FCB_ERR_INVALID, "element patch with %d accumulator rows (limit 254)", rows);
        max_rows = std::max(max_rows, rows);
        pacc_rows.push_back(rows);
        for (size_t j = 0; j < nodes.size(); ++j) {
            for (int kk = 0; kk < 4; ++kk) psrc.push_back(src[j][kk]);
            ++node_npatch[nodes[j]];
            uniq[nodes[j]] = -1;
        }
    }
    const int max_nodes = max_rows;
    // interior nodes are written directly; shared ones get a scratch slot per (patch, node)
    std::vector<std::vector<int>> slots_of(nN);
    int nslots = 0;
    for (int q = 0; q < npatch; ++q) {
        for (int nd : patch_nodes[q]) {
            if (node_npatch[nd] == 1) pnode_dst.push_back(nd);
            else { pnode_dst.push_back(-1 - nslots); slots_of[nd].push_back(nslots); ++nslots; }
        }
        pnode_ptr.push_back((int)pnode_dst.size());
    }
    auto rows_of = [&](int nd, std::vector<int>& out) {  // solver rows of the node's ux, uy and (vertex nodes) p dofs
        out.push_back(iperm[nd]);
        out.push_back(iperm[nd + nN]);
        out.push_back(nd < p->nV ? iperm[2 * nN + nd] : -1);
    };
    std::vector<int> prow, mrow;
    for (int q = 0; q < npatch; ++q)
        for (int nd : patch_nodes[q]) rows_of(nd, prow);
    std::vector<int> mptr(1, 0), msrc, mnode;
    for (int nd = 0; nd < nN; ++nd)
        if (node_npatch[nd] > 1) {
            for (int sl : slots_of[nd]) msrc.push_back(sl);
            mptr.push_back((int)msrc.size());
            mnode.push_back(nd);
            rows_of(nd, mrow);
        } else if (node_npatch[nd] == 0)
            return fail(h, FCB_ERR_INVALID, "P2 node %d belongs to no cell", nd);
    h->npatch = npatch;
    h->nshared = (int)mnode.size();
    int max_nn = 0;
    for (int q = 0; q < npatch; ++q) max_nn = std::max(max_nn, pnode_ptr[q + 1] - pnode_ptr[q]);
    h->patch_tab_off = max_nodes * 4 * 32 * (int)sizeof(double);
    h->patch_smem = h->patch_tab_off + ((5 * max_nn * (int)sizeof(int) + 15) & ~15);  // accumulators + node tables
    h->nblk_total = npatch;
    if (h->patch_smem > 220 * 1024) return fail(h, FCB_ERR_INVALID, "element patches too large for shared memory");
    CK(cudaFuncSetAttribute(k_element_patch<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->patch_smem));
    CK(cudaFuncSetAttribute(k_element_patch<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, h->patch_smem));
    TRY(upload(h, &h->pcell_ptr, pcell_ptr.data(), pcell_ptr.size()));
    TRY(upload(h, &h->pcnode, pcnode.data(), pcnode.size()));
    TRY(upload(h, &h->pgeo, pgeo.data(), pgeo.size()));
    TRY(upload(h, &h->plnode, plnode.data(), plnode.size()));
    TRY(upload(h, &h->pnode_ptr, pnode_ptr.data(), pnode_ptr.size()));
    TRY(upload(h, &h->pnode_dst, pnode_dst.data(), pnode_dst.size()));
    TRY(upload(h, &h->mptr, mptr.data(), mptr.size()));
    TRY(upload(h, &h->msrc, msrc.data(), std::max<size_t>(msrc.size(), 1)));
    TRY(upload(h, &h->mnode, mnode.data(), std::max<size_t>(mnode.size(), 1)));
    TRY(upload(h, &h->prow, prow.data(), std::max<size_t>(prow.size(), 1)));
    TRY(upload(h, &h->psrc, psrc.data(), std::max<size_t>(psrc.size(), 4)));
    TRY(upload(h, &h->pacc_rows, pacc_rows.data(), pacc_rows.size()));
    TRY(upload(h, &h->mrow, mrow.data(), std::max<size_t>(mrow.size(), 3)));
    TRY(upload<double>(h, &h->pscratch, nullptr, (size_t)std::max(nslots, 1) * 4 * h->ldb));
    CK(cudaStreamSynchronize(h->stream));  // host vectors go out of scope
    if (getenv("FCB_VERBOSE"))
        fprintf(stderr, "[fcb200] element patches: %d of <=%d cells, max %d accumulator rows (%d B smem), %d shared nodes in %d scratch slots\n",
                npatch, pc, max_rows, h->patch_smem, h->nshared, nslots);
    return FCB_OK;
}

int create_impl(fcb_context* h, const fcb_problem* p, int B) {
    CK(cudaSetDevice(h->device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, h->device));
    if (prop.major < 10) return fail(h, FCB_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", h->device, prop.major, prop.minor);
    h->num_sms = prop.multiProcessorCount;
    h->B = B;
    h->ldb = B <= 32 ? 32 : (B <= 64 ? 64 : (B + 127) / 128 * 128);  // a multiple of every sweep CTA width in use
    h->smem_per_sm = (int)prop.sharedMemPerMultiprocessor;
    {
        const int optin = (int)prop.sharedMemPerBlockOptin;
        CK(cudaFuncSetAttribute(k_front_sweep<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
        CK(cudaFuncSetAttribute(k_front_sweep<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
        CK(cudaFuncSetAttribute(k_front_sweep<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
        CK(cudaFuncSetAttribute(k_front_sweep<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
        CK(cudaFuncSetAttribute(k_front_sweep<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin));
        const char* env = getenv("FCB_SWEEP_WARPS");  // tuning knobs: force / cap the CTA width of the sweep launches
        h->force_nwc = env ? atoi(env) : 0;
        if (h->force_nwc != 1 && h->force_nwc != 2 && h->force_nwc != 4) h->force_nwc = 0;
        env = getenv("FCB_SWEEP_PDL");
        if (env) h->use_pdl = atoi(env) != 0;
        env = getenv("FCB_SWEEP_KSPLIT");
        if (env) h->allow_ksplit = atoi(env) != 0;
        env = getenv("FCB_SWEEP_SLOTS");
        if (env) {
            int v[4];
            if (sscanf(env, "%d,%d,%d,%d", &v[0], &v[1], &v[2], &v[3]) == 4)
                for (int i = 0; i < 4; ++i)
                    if (v[i] >= 12 && v[i] <= 96 && v[i] % 12 == 0) h->force_slots[i] = v[i];
        }
        env = getenv("FCB_SWEEP_WANT");
        if (env && atof(env) > 0.0) h->want_ctas_per_sm = atof(env);
        env = getenv("FCB_SWEEP_WANT_FWD");
        if (env && atof(env) > 0.0) h->want_ctas_fwd = atof(env);
        env = getenv("FCB_SWEEP_KSLOTS");
        if (env && atoi(env) >= 12 && atoi(env) <= 96 && atoi(env) % 12 == 0) h->kslots = atoi(env);
        env = getenv("FCB_SWEEP_MAXWARPS");
        if (env && (atoi(env) == 1 || atoi(env) == 2 || atoi(env) == 4)) h->max_nwc = atoi(env);  // 8 would need a 260-column box
        if (getenv("FCB_CLUSTER_DEBUG")) {
            CK(cudaMalloc((void**)&h->cl_dbg, 16 * DBG_PER_LAUNCH * sizeof(unsigned long long)));
            CK(cudaMemset(h->cl_dbg, 0, 16 * DBG_PER_LAUNCH * sizeof(unsigned long long)));
        }
        if (getenv("FCB_SWEEP_DEBUG")) {
            size_t nl = (size_t)std::max(p->plan[0].nlaunch, p->plan[1].nlaunch);
            CK(cudaMalloc((void**)&h->sweep_dbg, nl * DBG_PER_LAUNCH * sizeof(unsigned long long)));
            CK(cudaMemset(h->sweep_dbg, 0, nl * DBG_PER_LAUNCH * sizeof(unsigned long long)));
        }
        env = getenv("FCB_SWEEP_ROWBLOCKS");  // tuning knob: force the tile height (8-row blocks) of every launch
        h->force_nrb = env ? atoi(env) : 0;
        if (h->force_nrb < 0 || h->force_nrb > 4) h->force_nrb = 0;
    }
    CK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    for (auto& e : h->ev) CK(cudaEventCreate(&e));
    h->nT = p->nT; h->nN = p->nN; h->nV = p->nV;
    h->Nv = 2 * p->nN;
    h->N = h->Nv + p->nV;
    h->n = p->n_free; h->nbc = p->n_bc; h->na = p->na; h->ns = p->ns;
    h->dt = p->dt; h->nonlinear = p->nonlinear;
    if (h->n + h->nbc != h->N) return fail(h, FCB_ERR_INVALID, "n_free (%d) + n_bc (%d) != N (%d)", h->n, h->nbc, h->N);
    if (p->plan[0].n != h->n || p->plan[1].n != h->n) return fail(h, FCB_ERR_INVALID, "plan size does not match n_free");
    if (p->plan[0].nU != p->plan[1].nU) return fail(h, FCB_ERR_INVALID, "the two plans must share one symbolic structure (nU differs)");
    if (!(p->dt > 0)) return fail(h, FCB_ERR_INVALID, "dt must be positive");
    {
        double phi[7][6], dphi[7][6][2], w[7], mass[6][6];
        fill_tables(phi, dphi, w, mass);
        CK(cudaMemcpyToSymbol(c_phi, phi, sizeof phi));
        CK(cudaMemcpyToSymbol(c_dphi, dphi, sizeof dphi));
        CK(cudaMemcpyToSymbol(c_w, w, sizeof w));
        CK(cudaMemcpyToSymbol(c_mass, mass, sizeof mass));
    }
    TRY(upload(h, &h->cell_nodes, p->cell_nodes, (size_t)p->nT * 6));
    TRY(upload(h, &h->Jinv, p->Jinv, (size_t)p->nT * 4));
    TRY(upload(h, &h->detJ, p->detJ, (size_t)p->nT));
    for (size_t k = 0; k < (size_t)p->nT * 6; ++k)
        if (p->cell_nodes[k] < 0 || p->cell_nodes[k] >= p->nN) return fail(h, FCB_ERR_INVALID, "cell_nodes out of range");
    TRY(upload(h, &h->perm, p->perm, (size_t)h->n));
    {
        std::vector<int> iperm(h->N, 0);
        std::vector<char> seen(h->N, 0);
        for (int r = 0; r < h->n; ++r) {
            const int d = p->perm[r];
            if (d < 0 || d >= h->N || seen[d]) return fail(h, FCB_ERR_INVALID, "perm is not a valid injection");
            seen[d] = 1;
            iperm[d] = r;
        }
        for (int j = 0; j < h->nbc; ++j) {
            const int d = p->bc_dofs[j];
            if (d < 0 || d >= h->N || seen[d]) return fail(h, FCB_ERR_INVALID, "bc_dofs overlaps perm or is out of range");
            seen[d] = 1;
            iperm[d] = -1 - j;
        }
        TRY(upload(h, &h->iperm, iperm.data(), iperm.size()));
        CK(cudaStreamSynchronize(h->stream));
        TRY(build_patches(h, p, iperm));
    }
    TRY(upload(h, &h->bc_shape, p->bc_shape, (size_t)p->na * p->n_bc));
    for (int o = 0; o < 2; ++o) TRY(upload(h, &h->ctrl_rhs[o], p->ctrl_rhs[o], (size_t)p->na * p->n_free));
    TRY(upload(h, &h->bc_dofs, p->bc_dofs, (size_t)p->n_bc));
    // the control terms of the steady-state right-hand side touch only the rows next to the actuated boundaries / force support
    auto sparse_rows = [&](const double* dense, int* count, int** drows, double** dcoef) -> int {
        std::vector<int> rows;
        for (int r = 0; r < p->n_free; ++r)
            for (int k = 0; k < p->na; ++k)
                if (dense[(size_t)k * p->n_free + r] != 0.0) { rows.push_back(r); break; }
        std::vector<double> coef((size_t)p->na * rows.size());
        for (int k = 0; k < p->na; ++k)
            for (size_t i = 0; i < rows.size(); ++i) coef[(size_t)k * rows.size() + i] = dense[(size_t)k * p->n_free + rows[i]];
        *count = (int)rows.size();
        TRY(upload(h, drows, rows.data(), std::max<size_t>(rows.size(), 1)));
        TRY(upload(h, dcoef, coef.data(), std::max<size_t>(coef.size(), 1)));
        CK(cudaStreamSynchronize(h->stream));
        return FCB_OK;
    };
    TRY(sparse_rows(p->ctrl_rhs[1], &h->ncrow, &h->crow, &h->ccoef));
    h->scheme = p->scheme;
    if (p->scheme == 1) {
        if (!p->cn_ptr || !p->cn_idx || !p->cn_val || (p->na > 0 && !p->ctrl_rhs_prev))
            return fail(h, FCB_ERR_INVALID, "scheme 1 (Crank-Nicolson) needs cn_ptr/cn_idx/cn_val and ctrl_rhs_prev");
        const size_t nnz = (size_t)p->cn_ptr[p->n_free];
        for (size_t j = 0; j < nnz; ++j)
            if (p->cn_idx[j] < 0 || p->cn_idx[j] >= h->Nv) return fail(h, FCB_ERR_INVALID, "cn_idx out of the velocity range");
        TRY(upload(h, &h->cn_ptr, p->cn_ptr, (size_t)p->n_free + 1));
        TRY(upload(h, &h->cn_idx, p->cn_idx, std::max<size_t>(nnz, 1)));
        TRY(upload(h, &h->cn_val, p->cn_val, std::max<size_t>(nnz, 1)));
        if (p->na > 0) TRY(sparse_rows(p->ctrl_rhs_prev, &h->ncrow_prev, &h->crow_prev, &h->ccoef_prev));
        h->sp_csr = getenv("FCB_SPMM_CSR") && atoi(getenv("FCB_SPMM_CSR")) != 0;
        TRY(build_spmm(h, p));
    } else if (p->scheme != 0) {
        return fail(h, FCB_ERR_INVALID, "scheme must be 0 (BDF) or 1 (Crank-Nicolson)");
    }
    if (p->ns < 0 || (p->ns > 0 && (!p->sensor_ptr || p->sensor_ptr[0] != 0))) return fail(h, FCB_ERR_INVALID, "sensor_ptr must start at 0");
    for (int i = 0; i < p->ns; ++i) {
        if (p->sensor_ptr[i + 1] < p->sensor_ptr[i]) return fail(h, FCB_ERR_INVALID, "sensor_ptr must be non-decreasing (row %d)", i);
        for (int j = p->sensor_ptr[i]; j < p->sensor_ptr[i + 1]; ++j)
            if (p->sensor_idx[j] < 0 || p->sensor_idx[j] >= h->N) return fail(h, FCB_ERR_INVALID, "sensor_idx[%d] = %d is outside [0, %d)", j, p->sensor_idx[j], h->N);
    }
    TRY(upload(h, &h->sensor_ptr, p->sensor_ptr, (size_t)p->ns + 1));
    const size_t snnz = p->ns ? (size_t)p->sensor_ptr[p->ns] : 0;
    TRY(upload(h, &h->sensor_idx, p->sensor_idx, snnz));
    TRY(upload(h, &h->sensor_val, p->sensor_val, snnz));
    for (int o = 0; o < 2; ++o) {
        TRY(upload_plan(h, h->plan[o], p->plan[o], p->perm));
        TRY(upload_clusters(h, h->plan[o], p->plan[o], p->perm));
    }
    CK(cudaFuncSetAttribute(k_cluster_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prop.sharedMemPerBlockOptin));
    const size_t L = (size_t)h->ldb;
    for (int i = 0; i < 2; ++i) {
        TRY(upload<double>(h, &h->up[i], nullptr, (size_t)h->N * L));
        TRY(upload<double>(h, &h->bvec[i], nullptr, (size_t)h->Nv * L));
    }
    TRY(upload<double>(h, &h->avec, nullptr, (size_t)h->Nv * L));
    {
        const size_t zrows = (size_t)2 * h->n + (size_t)p->plan[0].nU;
        TRY(upload<double>(h, &h->Z, nullptr, (zrows + 1) * L));
        // TMA view of Z: [zrows][ldb] doubles; boxes of 1/2/4/8 rows x 32/64/128/256 columns.  Rows >= zrows are
        // outside the tensor and read as zeros, which is what the plan's "absent" gather index (2n + nU) needs.
        typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) return fail(h, FCB_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
        for (int wi = 0; wi < 4; ++wi)
            for (int hi = 0; hi < 6; ++hi) {
                const cuuint32_t width = 32u << wi;
                if ((int)width > h->ldb || width + 4 > 256) { memset(&h->zmaps[wi].m[hi], 0, sizeof(CUtensorMap)); continue; }
                const cuuint64_t gdim[2] = {(cuuint64_t)h->ldb, (cuuint64_t)zrows};
                const cuuint64_t gstride[1] = {(cuuint64_t)h->ldb * sizeof(double)};
                const cuuint32_t box[2] = {width + 4, 1u << hi};  // 4 spare columns: shared-memory row stride W+4 (see k_front_sweep)
                const cuuint32_t estride[2] = {1, 1};
                const CUresult r = ((EncodeTiled)fn)(&h->zmaps[wi].m[hi], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, h->Z, gdim, gstride, box,
                                                     estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                if (r != CUDA_SUCCESS) return fail(h, FCB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for a %ux%u box", (int)r, width, 1u << hi);
            }
    }
    TRY(upload<double>(h, &h->epart, nullptr, (size_t)h->nblk_total * L));
    TRY(upload<double>(h, &h->uctrl, nullptr, (size_t)(h->na > 0 ? h->na : 1) * L));
    TRY(upload<double>(h, &h->uctrl_prev, nullptr, (size_t)(h->na > 0 ? h->na : 1) * L));
    TRY(upload<double>(h, &h->y, nullptr, ((size_t)h->ns + 1) * L));  // y rows, then the dE row: one copy brings both back
    h->dE = h->y + (size_t)h->ns * L;
    TRY(upload<int>(h, &h->diverged, nullptr, L));
    TRY(upload<RunCtl>(h, &h->ctl, nullptr, 1));
    {
        const char* env = getenv("FCB_SERIES_CHUNK");
        if (env && atoi(env) >= 1 && atoi(env) <= (1 << 20)) h->series_chunk = atoi(env);
    }
    TRY(upload<double>(h, &h->costs, nullptr, 3 * L));
    if (cudaMallocHost((void**)&h->pin_in, (size_t)std::max(h->na, 1) * h->B * sizeof(double)) != cudaSuccess ||
        cudaMallocHost((void**)&h->pin_out, ((size_t)h->ns + 2) * h->B * sizeof(double)) != cudaSuccess) {
        cudaGetLastError();  // no pinned memory: the step falls back to direct (blocking) copies
        if (h->pin_in) cudaFreeHost(h->pin_in);
        h->pin_in = h->pin_out = nullptr;
    }
    CK(cudaStreamSynchronize(h->stream));
    return FCB_OK;
}

}  // namespace

// ----------------------------------------------------------------------------------------------
// C ABI
// ----------------------------------------------------------------------------------------------
extern "C" {

const char* fcb_version(void) { return "fcb200 0.1 sm_100a"; }

const char* fcb_last_error(fcb_handle h) { return h ? h->error.c_str() : g_create_error.c_str(); }

int fcb_create(const fcb_problem* problem, int32_t B, int32_t device, fcb_handle* out) {
    if (!problem || !out || B <= 0) return fail(nullptr, FCB_ERR_INVALID, "fcb_create: bad arguments");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(nullptr, FCB_ERR_NO_DEVICE, "fcb_create: no CUDA device available (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(nullptr, FCB_ERR_INVALID, "fcb_create: device %d out of range [0,%d)", device, ndev);
    fcb_context* h = new fcb_context();
    h->device = device;
    int rc = create_impl(h, problem, B);
    if (rc != FCB_OK) {
        g_create_error = h->error;
        destroy(h);
        *out = nullptr;
        return rc;
    }
    *out = h;
    return FCB_OK;
}

int fcb_destroy(fcb_handle h) {
    destroy(h);
    return FCB_OK;
}

int fcb_set_state(fcb_handle h, const double* u_n, const double* u_nn, const double* p_n, int32_t order) {
    if (!h || !u_n) return fail(h, FCB_ERR_INVALID, "fcb_set_state: null argument");
    if (order != 1 && order != 2) return fail(h, FCB_ERR_INVALID, "fcb_set_state: order must be 1 or 2");
    CK(cudaSetDevice(h->device));
    const size_t L = (size_t)h->ldb;
    for (int i = 0; i < 2; ++i) CK(cudaMemsetAsync(h->up[i], 0, (size_t)h->N * L * sizeof(double), h->stream));
    CK(cudaMemsetAsync(h->diverged, 0, L * sizeof(int), h->stream));
    h->parity = 0;
    TRY(copy_in(h, h->up[0], u_n, h->Nv));
    TRY(copy_in(h, h->up[1], u_nn ? u_nn : u_n, h->Nv));
    if (p_n) TRY(copy_in(h, h->up[0] + (size_t)h->Nv * L, p_n, h->nV));
    if (h->scheme == 1) {
        // Crank-Nicolson is self-starting: rhs = E u_n - N(u_n) + control terms, u_ctrl^{n} = 0 before the first step
        CK(cudaMemsetAsync(h->uctrl_prev, 0, (size_t)(h->na > 0 ? h->na : 1) * L * sizeof(double), h->stream));
        TRY(enqueue_spmm(h, h->up[0]));
        TRY(enqueue_element(h, h->up[0], h->avec, h->bvec[0], nullptr));
        h->rhs_ready = true;
        TRY(enqueue_measure(h, h->up[0]));
        CK(cudaStreamSynchronize(h->stream));
        h->order = 2;
        h->have_state = true;
        return FCB_OK;
    }
    // b_{n-1} from u_nn, then (a_n, b_n) and the energy partials from u_n
    TRY(enqueue_element(h, h->up[1], h->avec, h->bvec[1], nullptr));
    // a BDF2 (re)start goes through the same fused arithmetic as a running simulation, so that it continues a run
    // bit for bit; a BDF1 start needs a(u_n) itself (rhs = a/2)
    TRY(enqueue_element(h, h->up[0], h->avec, h->bvec[0], order == 2 ? h->bvec[1] : nullptr));
    h->rhs_ready = (order == 2);
    TRY(enqueue_measure(h, h->up[0]));
    CK(cudaStreamSynchronize(h->stream));
    h->order = order;
    h->have_state = true;
    return FCB_OK;
}

int fcb_set_controllers(fcb_handle h, const fcb_controllers* c) {
    if (!h || !c) return fail(h, FCB_ERR_INVALID, "fcb_set_controllers: null argument");
    if (c->nx < 0 || c->ny < 1 || c->nu < 1 || c->ny > 8 || c->nu > 8)
        return fail(h, FCB_ERR_INVALID, "fcb_set_controllers: need 1 <= ny,nu <= 8");
    CK(cudaSetDevice(h->device));
    void* old[] = {h->Ad, h->Bd, h->Cd, h->Dd, h->Ky, h->Fu, h->xk[0], h->xk[1]};
    for (void* p : old)
        if (p) cudaFree(p);
    h->Ad = h->Bd = h->Cd = h->Dd = h->Ky = h->Fu = h->xk[0] = h->xk[1] = nullptr;
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j)
            if (h->g_loop[i][j]) { cudaGraphExecDestroy(h->g_loop[i][j]); h->g_loop[i][j] = nullptr; }
    h->nx = c->nx; h->ny = c->ny; h->nu = c->nu;
    const size_t L = (size_t)h->ldb;
    auto up2 = [&](double** dst, const double* src, int rows) -> int {
        TRY(upload<double>(h, dst, nullptr, (size_t)(rows > 0 ? rows : 1) * L));
        if (src && rows > 0) TRY(copy_in(h, *dst, src, rows));
        return FCB_OK;
    };
    TRY(up2(&h->Ad, c->Ad, c->nx * c->nx));
    TRY(up2(&h->Bd, c->Bd, c->nx * c->ny));
    TRY(up2(&h->Cd, c->Cd, c->nu * c->nx));
    TRY(up2(&h->Dd, c->Dd, c->nu * c->ny));
    TRY(up2(&h->xk[0], c->x0, c->nx));
    TRY(up2(&h->xk[1], nullptr, c->nx));
    TRY(upload(h, &h->Ky, c->Ky, (size_t)c->ny * h->ns));
    TRY(upload(h, &h->Fu, c->Fu, (size_t)h->na * c->nu));
    CK(cudaStreamSynchronize(h->stream));
    h->xparity = 0;
    h->have_ctrl = true;
    return FCB_OK;
}

// pageable host memory (plain numpy arrays): copies to and from it block the calling thread one by one; the step
// stages them through pinned buffers of the handle instead, so that all transfers are asynchronous and the call
// synchronises once
static bool is_pageable_host(const void* p) {
    // callers pass the same few buffers step after step: remember the last answers (a driver query costs microseconds)
    struct Seen { const void* p; bool pageable; };
    static thread_local Seen seen[8] = {};
    static thread_local int next = 0;
    for (const Seen& s : seen)
        if (s.p == p && p) return s.pageable;
    cudaPointerAttributes at;
    bool pageable = true;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) cudaGetLastError();
    else pageable = at.type == cudaMemoryTypeUnregistered;
    seen[next] = {p, pageable};
    next = (next + 1) % 8;
    return pageable;
}

int fcb_step(fcb_handle h, const double* u_ctrl, double* y_meas, double* dE, int32_t* diverged) {
    if (!h) return FCB_ERR_INVALID;
    if (!h->have_state) return fail(h, FCB_ERR_STATE, "fcb_step: call fcb_set_state first");
    if (h->na > 0 && !u_ctrl) return fail(h, FCB_ERR_INVALID, "fcb_step: u_ctrl is NULL");
    CK(cudaSetDevice(h->device));
    const size_t B = (size_t)h->B;
    if (h->na > 0) {
        if (h->pin_in && is_pageable_host(u_ctrl)) {
            memcpy(h->pin_in, u_ctrl, (size_t)h->na * B * sizeof(double));
            TRY(copy_in(h, h->uctrl, h->pin_in, h->na));
        } else {
            TRY(copy_in(h, h->uctrl, u_ctrl, h->na));
        }
    }
    TRY(run_one_step(h, RUN_STEP));
    double* py = h->pin_out;                                   // [ns][B] y, [B] dE, then [B] ints
    double* pe = h->pin_out ? h->pin_out + (size_t)h->ns * B : nullptr;
    int32_t* pd = h->pin_out ? reinterpret_cast<int32_t*>(pe + B) : nullptr;
    const bool sy = y_meas && h->ns > 0 && h->pin_out && is_pageable_host(y_meas);
    const bool se = dE && h->pin_out && is_pageable_host(dE);
    const bool sd = diverged && h->pin_out && is_pageable_host(diverged);
    if (sy && se) {
        TRY(copy_out(h, py, h->y, h->ns + 1));  // y rows and the dE row are adjacent on both sides: one copy
    } else {
        if (y_meas && h->ns > 0) TRY(copy_out(h, sy ? py : y_meas, h->y, h->ns));
        if (dE) TRY(copy_out(h, se ? pe : dE, h->dE, 1));
    }
    if (diverged) TRY(copy_out(h, sd ? pd : diverged, h->diverged, 1));
    CK(cudaStreamSynchronize(h->stream));
    if (sy) memcpy(y_meas, py, (size_t)h->ns * B * sizeof(double));
    if (se) memcpy(dE, pe, B * sizeof(double));
    if (sd) memcpy(diverged, pd, B * sizeof(int32_t));
    return FCB_OK;
}

// Device-resident run of nsteps steps (closed loop: controller -> step -> log; open loop: load u_ctrl -> step -> log), streamed
// chunk by chunk: the handle owns one series buffer of `series_chunk` steps; after every chunk its rows go out to the
// caller's array (asynchronously, in stream order) and the next chunk replays the SAME graphs -- the kernels read the
// buffer, its capacity and the step counter from the RunCtl block, so neither the run length nor logging on/off ever
// forces a re-capture, and a long run needs no series-sized device allocation.
static int run_device_loop(fcb_handle h, int mode, int32_t nsteps, const double* u_series, double* series) {
    CK(cudaSetDevice(h->device));
    const int ncol = 1 + h->na + h->ns;
    const int chunk = std::max(1, std::min(h->series_chunk, (int)nsteps));
    const bool want_u = mode == RUN_OPEN && h->na > 0;
    if (chunk > h->series_capacity) {
        CK(cudaStreamSynchronize(h->stream));
        if (h->series) cudaFree(h->series);
        if (h->useries) cudaFree(h->useries);
        h->series = h->useries = nullptr;
        h->series_capacity = 0;
        CK(cudaMalloc((void**)&h->series, (size_t)chunk * ncol * h->ldb * sizeof(double)));
        CK(cudaMalloc((void**)&h->useries, (size_t)chunk * std::max(h->na, 1) * h->ldb * sizeof(double)));
        CK(cudaMemsetAsync(h->useries, 0, (size_t)chunk * std::max(h->na, 1) * h->ldb * sizeof(double), h->stream));
        h->series_capacity = chunk;
    }
    CK(cudaMemsetAsync(h->costs, 0, 3 * (size_t)h->ldb * sizeof(double), h->stream));
    for (int c0 = 0; c0 < nsteps; c0 += chunk) {
        const int nc = std::min(chunk, (int)nsteps - c0);
        const RunCtl ctl{h->series, h->useries, series ? nc : 0, 0};
        CK(cudaMemcpyAsync(h->ctl, &ctl, sizeof ctl, cudaMemcpyHostToDevice, h->stream));  // pageable source: staged before the call returns
        if (want_u)
            CK(cudaMemcpy2DAsync(h->useries, (size_t)h->ldb * sizeof(double), u_series + (size_t)c0 * h->na * h->B, (size_t)h->B * sizeof(double),
                                 (size_t)h->B * sizeof(double), (size_t)nc * h->na, cudaMemcpyDefault, h->stream));
        for (int s = 0; s < nc; ++s) TRY(run_one_step(h, mode));
        if (series)
            CK(cudaMemcpy2DAsync(series + (size_t)c0 * ncol * h->B, (size_t)h->B * sizeof(double), h->series, (size_t)h->ldb * sizeof(double),
                                 (size_t)h->B * sizeof(double), (size_t)nc * ncol, cudaMemcpyDefault, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    return FCB_OK;
}

int fcb_run_closed_loop(fcb_handle h, int32_t nsteps, double* series) {
    if (!h || nsteps < 0) return FCB_ERR_INVALID;
    if (!h->have_state) return fail(h, FCB_ERR_STATE, "fcb_run_closed_loop: call fcb_set_state first");
    if (!h->have_ctrl) return fail(h, FCB_ERR_STATE, "fcb_run_closed_loop: call fcb_set_controllers first");
    return run_device_loop(h, RUN_CLOSED, nsteps, nullptr, series);
}

int fcb_run_open_loop(fcb_handle h, int32_t nsteps, const double* u_series, double* series) {
    if (!h || nsteps < 0) return FCB_ERR_INVALID;
    if (!h->have_state) return fail(h, FCB_ERR_STATE, "fcb_run_open_loop: call fcb_set_state first");
    if (h->na > 0 && !u_series) return fail(h, FCB_ERR_INVALID, "fcb_run_open_loop: u_series is NULL");
    return run_device_loop(h, RUN_OPEN, nsteps, u_series, series);
}

int fcb_set_controller_state(fcb_handle h, const double* x) {
    if (!h) return FCB_ERR_INVALID;
    if (!h->have_ctrl) return fail(h, FCB_ERR_STATE, "fcb_set_controller_state: no controllers set");
    CK(cudaSetDevice(h->device));
    if (h->nx > 0) {
        if (x) TRY(copy_in(h, h->xk[h->xparity], x, h->nx));
        else CK(cudaMemsetAsync(h->xk[h->xparity], 0, (size_t)h->nx * h->ldb * sizeof(double), h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    return FCB_OK;
}

int fcb_get_fields(fcb_handle h, int32_t which, double* up) {
    if (!h || !up) return FCB_ERR_INVALID;
    if (!h->have_state) return fail(h, FCB_ERR_STATE, "fcb_get_fields: no state");
    CK(cudaSetDevice(h->device));
    const double* src = (which == 0) ? h->up[h->parity] : h->up[1 - h->parity];
    TRY(copy_out(h, up, src, which == 0 ? h->N : h->Nv));
    CK(cudaStreamSynchronize(h->stream));
    return FCB_OK;
}

int fcb_get_costs(fcb_handle h, double* costs) {
    if (!h || !costs) return FCB_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    TRY(copy_out(h, costs, h->costs, 3));
    CK(cudaStreamSynchronize(h->stream));
    return FCB_OK;
}

int fcb_get_measurement(fcb_handle h, double* y_meas, double* dE, int32_t* diverged) {
    if (!h) return FCB_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    if (y_meas && h->ns > 0) TRY(copy_out(h, y_meas, h->y, h->ns));
    if (dE) TRY(copy_out(h, dE, h->dE, 1));
    if (diverged) TRY(copy_out(h, diverged, h->diverged, 1));
    CK(cudaStreamSynchronize(h->stream));
    return FCB_OK;
}

int fcb_get_controller_state(fcb_handle h, double* x) {
    if (!h || !x) return FCB_ERR_INVALID;
    if (!h->have_ctrl) return fail(h, FCB_ERR_STATE, "no controllers set");
    CK(cudaSetDevice(h->device));
    if (h->nx > 0) TRY(copy_out(h, x, h->xk[h->xparity], h->nx));
    CK(cudaStreamSynchronize(h->stream));
    return FCB_OK;
}

int fcb_profile_step(fcb_handle h, const double* u_ctrl, float* ms, int32_t* launches) {
    if (!h || !ms) return FCB_ERR_INVALID;
    if (!h->have_state) return fail(h, FCB_ERR_STATE, "fcb_profile_step: call fcb_set_state first");
    CK(cudaSetDevice(h->device));
    if (h->na > 0 && u_ctrl) TRY(copy_in(h, h->uctrl, u_ctrl, h->na));
    PhaseMark pm{h, true};
    const DevPlan& pl = h->plan[h->order - 1];
    TRY(enqueue_step(h, h->order, h->parity, h->rhs_ready, &pm));
    CK(cudaStreamSynchronize(h->stream));
    h->parity ^= 1;
    h->order = 2;
    h->rhs_ready = true;
    for (int i = 0; i < FCB_NPHASES; ++i) CK(cudaEventElapsedTime(&ms[i], h->ev[i], h->ev[i + 1]));
    if (h->cl_dbg) {
        std::vector<unsigned long long> host(16 * DBG_PER_LAUNCH);
        CK(cudaMemcpy(host.data(), h->cl_dbg, host.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        if (FILE* f = fopen(getenv("FCB_CLUSTER_DEBUG"), "wb")) {
            const int nt = (int)std::min<size_t>(pl.tiers.size(), 8);
            fwrite(&nt, sizeof(int), 1, f);
            for (int t = 0; t < nt; ++t)
                for (int dir = 0; dir < 2; ++dir) {
                    const int nc = std::min(pl.tiers[t].count * (h->ldb / CL_W), 8192);
                    fwrite(&nc, sizeof(int), 1, f);
                    fwrite(host.data() + (size_t)(t * 2 + dir) * DBG_PER_LAUNCH, sizeof(unsigned long long), (size_t)nc * 8, f);
                }
            fclose(f);
        }
    }
    if (h->sweep_dbg) {
        const char* path = getenv("FCB_SWEEP_DEBUG");
        std::vector<unsigned long long> host((size_t)pl.nlaunch * DBG_PER_LAUNCH);
        CK(cudaMemcpy(host.data(), h->sweep_dbg, host.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        if (FILE* f = fopen(path, "wb")) {
            const int hdr[2] = {pl.nlaunch, pl.n_forward};
            fwrite(hdr, sizeof(int), 2, f);
            for (int l = 0; l < pl.nlaunch; ++l) {
                const DevPlan::Launch& L = pl.launches[l];
                const int meta[4] = {L.grid, L.nslab, L.nwc, L.nstages};
                fwrite(meta, sizeof(int), 4, f);
                fwrite(host.data() + (size_t)l * DBG_PER_LAUNCH, sizeof(unsigned long long), (size_t)L.grid * L.nslab * 8, f);
            }
            fclose(f);
        }
    }
    if (launches) {
        launches[FCB_PHASE_RHS] = 1 + (h->scheme == 1 && h->ncrow_prev > 0 ? 1 : 0);
        launches[FCB_PHASE_FORWARD] = pl.n_forward + (int)pl.tiers.size();
        int nsum_f = 0, nsum_b = 0;
        for (int l = 0; l < pl.nlaunch; ++l)
            if (pl.asm_lptr[l + 1] > pl.asm_lptr[l]) (l < pl.n_forward ? nsum_f : nsum_b) += 1;
        launches[FCB_PHASE_FORWARD] += nsum_f;
        launches[FCB_PHASE_BACKWARD] = pl.nlaunch - pl.n_forward + nsum_b + (int)pl.tiers.size();
        launches[FCB_PHASE_POST] = h->nbc > 0 ? 1 : 0;
        launches[FCB_PHASE_SPMM] = h->scheme == 1 ? 1 : 0;
        launches[FCB_PHASE_ELEMENT] = 1 + (h->nshared > 0 ? 1 : 0);
        launches[FCB_PHASE_MEASURE] = 1;
    }
    return FCB_OK;
}


int fcb_assemble_advection(const fcb_assembly* m, int32_t B, int32_t device, const double* U, double* C, double* D) {
    fcb_context* h = nullptr;  // errors go to fcb_last_error(NULL)
    if (!m || !U || !C || B <= 0) return fail(h, FCB_ERR_INVALID, "fcb_assemble_advection: bad arguments");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(h, FCB_ERR_NO_DEVICE, "fcb_assemble_advection: no CUDA device available (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(h, FCB_ERR_INVALID, "fcb_assemble_advection: device %d out of range", device);
    if (m->nT <= 0 || m->nN <= 0 || m->nnz <= 0 || m->ncolour <= 0 || !m->cell_nodes || !m->Jinv || !m->detJ || !m->colour_ptr || !m->colour_cells || !m->pos)
        return fail(h, FCB_ERR_INVALID, "fcb_assemble_advection: incomplete mesh description");
    if (m->colour_ptr[0] != 0 || m->colour_ptr[m->ncolour] != m->nT) return fail(h, FCB_ERR_INVALID, "fcb_assemble_advection: colour_ptr must cover every cell");
    {   // ranges, and the colouring contract: cells of one colour share no P2 node (otherwise the plain scatter would race)
        std::vector<int> seen((size_t)m->nN, -1);
        for (int c = 0; c < m->ncolour; ++c) {
            if (m->colour_ptr[c + 1] < m->colour_ptr[c]) return fail(h, FCB_ERR_INVALID, "fcb_assemble_advection: colour_ptr must be non-decreasing");
            for (int k = m->colour_ptr[c]; k < m->colour_ptr[c + 1]; ++k) {
                const int e = m->colour_cells[k];
                if (e < 0 || e >= m->nT) return fail(h, FCB_ERR_INVALID, "fcb_assemble_advection: colour_cells out of range");
                for (int i = 0; i < 6; ++i) {
                    const int nd = m->cell_nodes[(size_t)e * 6 + i];
                    if (nd < 0 || nd >= m->nN) return fail(h, FCB_ERR_INVALID, "fcb_assemble_advection: cell_nodes out of range");
                    if (seen[nd] == c) return fail(h, FCB_ERR_INVALID, "fcb_assemble_advection: two cells of colour %d share node %d", c, nd);
                    seen[nd] = c;
                }
            }
        }
        for (size_t k = 0; k < (size_t)m->nT * 36; ++k)
            if (m->pos[k] < 0 || m->pos[k] >= m->nnz) return fail(h, FCB_ERR_INVALID, "fcb_assemble_advection: pos out of range");
    }
#define CKA(call)                                                                                                        \
    do {                                                                                                                 \
        cudaError_t e_ = (call);                                                                                         \
        if (e_ != cudaSuccess) {                                                                                         \
            for (void* q_ : bufs) cudaFree(q_);                                                                          \
            return fail(h, FCB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);     \
        }                                                                                                                \
    } while (0)
    std::vector<void*> bufs;
    CKA(cudaSetDevice(device));
    {
        double phi[7][6], dphi[7][6][2], w[7], mass[6][6];
        fill_tables(phi, dphi, w, mass);
        CKA(cudaMemcpyToSymbol(c_phi, phi, sizeof phi));
        CKA(cudaMemcpyToSymbol(c_dphi, dphi, sizeof dphi));
        CKA(cudaMemcpyToSymbol(c_w, w, sizeof w));
        CKA(cudaMemcpyToSymbol(c_mass, mass, sizeof mass));
    }
    const int ldb = (B + 31) / 32 * 32;
    const size_t L = (size_t)ldb, Nv = 2 * (size_t)m->nN, nnz = (size_t)m->nnz;
    int *d_cells = nullptr, *d_nodes = nullptr, *d_pos = nullptr;
    double *d_J = nullptr, *d_det = nullptr, *d_U = nullptr, *d_C = nullptr, *d_D = nullptr;
    auto dev_alloc = [&](void** p, size_t bytes) { cudaError_t e = cudaMalloc(p, bytes); if (e == cudaSuccess) bufs.push_back(*p); return e; };
    CKA(dev_alloc((void**)&d_cells, (size_t)m->nT * sizeof(int)));
    CKA(dev_alloc((void**)&d_nodes, (size_t)m->nT * 6 * sizeof(int)));
    CKA(dev_alloc((void**)&d_pos, (size_t)m->nT * 36 * sizeof(int)));
    CKA(dev_alloc((void**)&d_J, (size_t)m->nT * 4 * sizeof(double)));
    CKA(dev_alloc((void**)&d_det, (size_t)m->nT * sizeof(double)));
    CKA(dev_alloc((void**)&d_U, Nv * L * sizeof(double)));
    CKA(dev_alloc((void**)&d_C, nnz * L * sizeof(double)));
    if (D) CKA(dev_alloc((void**)&d_D, 4 * nnz * L * sizeof(double)));
    CKA(cudaMemcpy(d_cells, m->colour_cells, (size_t)m->nT * sizeof(int), cudaMemcpyDefault));
    CKA(cudaMemcpy(d_nodes, m->cell_nodes, (size_t)m->nT * 6 * sizeof(int), cudaMemcpyDefault));
    CKA(cudaMemcpy(d_pos, m->pos, (size_t)m->nT * 36 * sizeof(int), cudaMemcpyDefault));
    CKA(cudaMemcpy(d_J, m->Jinv, (size_t)m->nT * 4 * sizeof(double), cudaMemcpyDefault));
    CKA(cudaMemcpy(d_det, m->detJ, (size_t)m->nT * sizeof(double), cudaMemcpyDefault));
    CKA(cudaMemset(d_U, 0, Nv * L * sizeof(double)));
    CKA(cudaMemcpy2D(d_U, L * sizeof(double), U, (size_t)B * sizeof(double), (size_t)B * sizeof(double), Nv, cudaMemcpyDefault));
    CKA(cudaMemset(d_C, 0, nnz * L * sizeof(double)));
    if (D) CKA(cudaMemset(d_D, 0, 4 * nnz * L * sizeof(double)));
    for (int c = 0; c < m->ncolour; ++c) {
        const int c0 = m->colour_ptr[c], nc = m->colour_ptr[c + 1] - c0;
        if (nc == 0) continue;
        k_assemble_advection<<<dim3((nc + 3) / 4, ldb / 32), dim3(32, 4)>>>(c0, nc, d_cells, d_nodes, d_J, d_det, d_pos, d_U, d_C, d_D, m->nN, nnz, ldb);
    }
    CKA(cudaGetLastError());
    CKA(cudaMemcpy2D(C, (size_t)B * sizeof(double), d_C, L * sizeof(double), (size_t)B * sizeof(double), nnz, cudaMemcpyDefault));
    if (D) CKA(cudaMemcpy2D(D, (size_t)B * sizeof(double), d_D, L * sizeof(double), (size_t)B * sizeof(double), 4 * nnz, cudaMemcpyDefault));
    CKA(cudaDeviceSynchronize());
    for (void* q_ : bufs) cudaFree(q_);
#undef CKA
    return FCB_OK;
}


int fcb_factorize(const fcb_symbolic* sy, int32_t device, const double* avals, int64_t nnz, double* E, double* Finv, double* G, double* growth) {
    fcb_context* h = nullptr;  // errors go to fcb_last_error(NULL)
    if (!sy || !avals || !E || !Finv || !G || !growth || nnz < 0) return fail(h, FCB_ERR_INVALID, "fcb_factorize: bad arguments");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(h, FCB_ERR_NO_DEVICE, "fcb_factorize: no CUDA device available (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(h, FCB_ERR_INVALID, "fcb_factorize: device %d out of range", device);
    const int nf = sy->nfront;
    if (nf <= 0 || sy->nlevel <= 0 || !sy->w || !sy->m || !sy->level_ptr || !sy->level_fronts || !sy->a_ptr || !sy->c_ptr || !sy->c_lptr)
        return fail(h, FCB_ERR_INVALID, "fcb_factorize: incomplete symbolic structure");
    if (sy->level_ptr[0] != 0 || sy->level_ptr[sy->nlevel] != nf) return fail(h, FCB_ERR_INVALID, "fcb_factorize: level_ptr must cover every front");
    std::vector<FrontDesc> fd((size_t)nf);
    long long foff = 0, eoff = 0, ioff = 0;
    int wmax = 0;
    for (int f = 0; f < nf; ++f) {
        const int w = sy->w[f], m = sy->m[f];
        if (w < 0 || m < 0) return fail(h, FCB_ERR_INVALID, "fcb_factorize: negative front size");
        fd[f] = {foff, eoff, ioff, w, m};
        foff += (long long)(w + m) * (w + m);
        eoff += (long long)m * w;
        ioff += (long long)w * w;
        wmax = std::max(wmax, w);
        for (long long k = sy->a_ptr[f]; k < sy->a_ptr[f + 1]; ++k)
            if (sy->a_src[k] < 0 || sy->a_src[k] >= nnz || sy->a_dst[k] < 0 || sy->a_dst[k] >= (long long)(w + m) * (w + m))
                return fail(h, FCB_ERR_INVALID, "fcb_factorize: matrix entry map of front %d out of range", f);
        for (long long c = sy->c_ptr[f]; c < sy->c_ptr[f + 1]; ++c) {
            const int ch = sy->c_front[c];
            if (ch < 0 || ch >= nf || sy->c_lptr[c + 1] - sy->c_lptr[c] != sy->m[ch]) return fail(h, FCB_ERR_INVALID, "fcb_factorize: child list of front %d is malformed", f);
            for (long long t = sy->c_lptr[c]; t < sy->c_lptr[c + 1]; ++t)
                if (sy->c_loc[t] < 0 || sy->c_loc[t] >= w + m) return fail(h, FCB_ERR_INVALID, "fcb_factorize: extend-add map of front %d out of range", f);
        }
    }
    {   // children must sit in earlier levels
        std::vector<int> level_of((size_t)nf, -1);
        for (int l = 0; l < sy->nlevel; ++l)
            for (int k = sy->level_ptr[l]; k < sy->level_ptr[l + 1]; ++k) {
                const int f = sy->level_fronts[k];
                if (f < 0 || f >= nf || level_of[f] >= 0) return fail(h, FCB_ERR_INVALID, "fcb_factorize: level_fronts is not a permutation");
                level_of[f] = l;
            }
        for (int f = 0; f < nf; ++f)
            for (long long c = sy->c_ptr[f]; c < sy->c_ptr[f + 1]; ++c)
                if (level_of[sy->c_front[c]] >= level_of[f]) return fail(h, FCB_ERR_INVALID, "fcb_factorize: a child of front %d is not in an earlier level", f);
    }
    std::vector<void*> bufs;
#define CKF(call)                                                                                                        \
    do {                                                                                                                 \
        cudaError_t e_ = (call);                                                                                         \
        if (e_ != cudaSuccess) {                                                                                         \
            for (void* q_ : bufs) cudaFree(q_);                                                                          \
            return fail(h, FCB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);     \
        }                                                                                                                \
    } while (0)
    CKF(cudaSetDevice(device));
    auto dev_copy = [&](void** p, const void* src, size_t bytes) {
        cudaError_t e = cudaMalloc(p, std::max<size_t>(bytes, 16));
        if (e != cudaSuccess) return e;
        bufs.push_back(*p);
        return src ? cudaMemcpy(*p, src, bytes, cudaMemcpyDefault) : cudaMemset(*p, 0, std::max<size_t>(bytes, 16));
    };
    const long long nch = sy->c_ptr[nf];
    int *d_fronts, *d_asrc, *d_adst, *d_cfront, *d_cloc;
    long long *d_aptr, *d_cptr, *d_clptr;
    FrontDesc* d_fd;
    double *d_av, *d_W, *d_E, *d_I, *d_G, *d_gr;
    CKF(dev_copy((void**)&d_fronts, sy->level_fronts, (size_t)nf * 4));
    CKF(dev_copy((void**)&d_fd, fd.data(), (size_t)nf * sizeof(FrontDesc)));
    CKF(dev_copy((void**)&d_aptr, sy->a_ptr, (size_t)(nf + 1) * 8));
    CKF(dev_copy((void**)&d_asrc, sy->a_src, (size_t)sy->a_ptr[nf] * 4));
    CKF(dev_copy((void**)&d_adst, sy->a_dst, (size_t)sy->a_ptr[nf] * 4));
    CKF(dev_copy((void**)&d_cptr, sy->c_ptr, (size_t)(nf + 1) * 8));
    CKF(dev_copy((void**)&d_cfront, sy->c_front, (size_t)nch * 4));
    CKF(dev_copy((void**)&d_clptr, sy->c_lptr, (size_t)(nch + 1) * 8));
    CKF(dev_copy((void**)&d_cloc, sy->c_loc, (size_t)sy->c_lptr[nch] * 4));
    CKF(dev_copy((void**)&d_av, avals, (size_t)nnz * 8));
    CKF(dev_copy((void**)&d_W, nullptr, (size_t)foff * 8));
    CKF(dev_copy((void**)&d_E, nullptr, (size_t)eoff * 8));
    CKF(dev_copy((void**)&d_G, nullptr, (size_t)eoff * 8));
    CKF(dev_copy((void**)&d_I, nullptr, (size_t)ioff * 8));
    CKF(dev_copy((void**)&d_gr, nullptr, (size_t)nf * 8));
    for (int l = 0; l < sy->nlevel; ++l) {
        const int k0 = sy->level_ptr[l], nl = sy->level_ptr[l + 1] - k0;
        if (nl <= 0) continue;
        int tiles = 1;
        for (int k = k0; k < k0 + nl; ++k) {
            const int f = sy->level_fronts[k], w = sy->w[f], m = sy->m[f];
            tiles = std::max(tiles, ((std::max(w, m) + 31) / 32) * ((std::max(w, m) + 31) / 32));
        }
        tiles = std::min(tiles, 1024);
        k_ff_assemble<<<nl, 512>>>(d_fronts + k0, d_fd, d_aptr, d_asrc, d_adst, d_av, d_cptr, d_cfront, d_clptr, d_cloc, d_W);
        k_ff_invert<<<nl, 512, (size_t)std::max(wmax, 1) * sizeof(int)>>>(d_fronts + k0, d_fd, d_W, d_I, d_gr);
        k_ff_gemm<<<dim3(nl, tiles), dim3(16, 16)>>>(0, d_fronts + k0, d_fd, d_W, d_E, d_G);
        k_ff_gemm<<<dim3(nl, tiles), dim3(16, 16)>>>(1, d_fronts + k0, d_fd, d_W, d_E, d_G);
        k_ff_gemm<<<dim3(nl, tiles), dim3(16, 16)>>>(2, d_fronts + k0, d_fd, d_W, d_E, d_G);
        CKF(cudaGetLastError());
    }
    CKF(cudaMemcpy(E, d_E, (size_t)eoff * 8, cudaMemcpyDefault));
    CKF(cudaMemcpy(G, d_G, (size_t)eoff * 8, cudaMemcpyDefault));
    CKF(cudaMemcpy(Finv, d_I, (size_t)ioff * 8, cudaMemcpyDefault));
    CKF(cudaMemcpy(growth, d_gr, (size_t)nf * 8, cudaMemcpyDefault));
    CKF(cudaDeviceSynchronize());
    for (void* q_ : bufs) cudaFree(q_);
#undef CKF
    return FCB_OK;
}

int64_t fcb_launch_count(fcb_handle h) { return h ? h->launches : 0; }
void* fcb_stream(fcb_handle h) { return h ? (void*)h->stream : nullptr; }
int fcb_synchronize(fcb_handle h) {
    if (!h) return FCB_ERR_INVALID;
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    return FCB_OK;
}

}  // extern "C"
