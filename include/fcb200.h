/*
 * fcb200.h — C ABI of the B200-native ensemble replacement for FlowControl's
 * time-stepping hot path.
 *
 * The reference (williamjussiau/FlowControl) is pure Python and has no FFI of
 * its own; its hot path crosses into third-party native code through dolfin's
 * pybind11 layer.  Each entry point below names the reference call site(s) it
 * replaces (paths relative to /root/reference).  INTEGRATION.md shows the
 * ctypes stub a maintainer would add to the reference to bind them.
 *
 * Conventions
 *   - plain C: opaque handle, pointers + sizes, no torch / Python types.
 *   - every function returns 0 on success, a negative fcb_status otherwise;
 *     fcb_last_error(h) (or fcb_last_error(NULL) for fcb_create failures)
 *     returns a human-readable message.
 *   - all floating-point data are IEEE binary64, all maps int32.
 *   - ensemble arrays are "dof-major, trajectory-innermost": X[row * B + b],
 *     b in [0,B).  Data pointers may be host (pageable or pinned) or device
 *     pointers; the library copies with cudaMemcpyDefault (UVA).
 *   - one handle = one GPU = one CUDA stream; a handle is not thread-safe,
 *     independent handles are independent.
 *   - there is NO CPU fallback: fcb_create fails if no sm_100 device is usable.
 */
#ifndef FCB200_H
#define FCB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fcb_context* fcb_handle;

typedef enum fcb_status {
    FCB_OK = 0,
    FCB_ERR_INVALID = -1,   /* bad argument / inconsistent sizes          */
    FCB_ERR_CUDA = -2,      /* CUDA runtime error (see fcb_last_error)     */
    FCB_ERR_NO_DEVICE = -3, /* no usable GPU                               */
    FCB_ERR_STATE = -4      /* call order violated (e.g. step before state)*/
} fcb_status;

/* Solve plan of one factorised LHS (flowcontrol_b200/multifrontal.py: SolvePlan): the dense blocks of a
 * multifrontal factorisation in application order.  Replaces the MUMPS factors held by dolfin.LUSolver
 * after set_operator (src/flowcontrol/flowsolver.py:694-697, 812-814).
 * The device buffer Z has rows [0,n) b -> x, [n,2n) y, [2n,2n+nU) update vectors, row 2n+nU = zeros.
 * Block q:  x_k = Z[i0[k]] (+ Z[i1[k]] + Z[i2[k]] if nsrc == 3), k < K;  if ystore >= 0: Z[ystore+k] = x_k;
 *           Z[out0+r] = sum_k V[r][k] x_k (+ Z[e0[r]] + Z[e1[r]] if eptr >= 0, -1 = absent), r < M.
 * Blocks of one launch are independent (no block reads a row another block of that launch writes). */
typedef struct fcb_plan {
    int32_t n;                 /* free unknowns                                                    */
    int32_t nU;                /* rows of the update-vector region                                 */
    int32_t nblocks;
    const int32_t* blk_K;      /* [nblocks] gathered input rows                                    */
    const int32_t* blk_M;      /* [nblocks] output rows (0 = store-only block)                     */
    const int32_t* blk_nsrc;   /* [nblocks] 1 or 3 gather lists                                    */
    const int32_t* blk_out0;   /* [nblocks] first output row in Z                                  */
    const int32_t* blk_ystore; /* [nblocks] first row to store the gathered x_k to, or -1; -2 on a backward block of solution
                                * rows: nobody gathers them, they are written in canonical numbering only (out0 still names them) */
    const int64_t* blk_iptr;   /* [nblocks] offsets into i0/i1/i2 (K entries per block)            */
    const int64_t* blk_vptr;   /* [nblocks] offsets into vals (row-major M x K per block)          */
    const int64_t* blk_eptr;   /* [nblocks] offsets into e0/e1 (M entries per block), or -1        */
    const int32_t *i0, *i1, *i2;
    const int32_t *e0, *e1;
    const double* vals;
    int32_t nlaunch;
    const int32_t* launch_ptr; /* [nlaunch+1] block ranges; blocks of a launch are independent      */
    int32_t n_forward_launches;
    /* gather-sums: Z[asm_dst[i]] = sum of Z[asm_src[j]], asm_ptr[i] <= j < asm_ptr[i+1], in that order.  Rows
     * [asm_lptr[l], asm_lptr[l+1]) run right before launch l: the right-hand side of the merged top of the elimination tree,
     * the "virtual" update vectors of fronts with more than two children (amalgamated levels), and the pre-summed
     * y_t = b_t + sum of the children's update rows of forward blocks that gather one plane (build_plan(presum_height)).
     * asm_lptr == NULL: every row runs before launch n_forward_launches.  asm_n may be 0 */
    int32_t asm_n;
    const int32_t* asm_ptr;  /* [asm_n+1] */
    const int32_t* asm_src;
    const int32_t* asm_dst;  /* [asm_n] */
    const int32_t* asm_lptr; /* [nlaunch+1] or NULL */
    /* Subtree clusters (multifrontal.py: choose_clusters): connected pieces of the lower elimination tree that ONE CTA
     * sweeps with all their unknowns resident in shared memory (k_cluster_sweep).  Inside a cluster the sweep is
     * right-looking on one resident vector S = [own rows of its fronts | boundary rows of its root]:
     *   forward   S[own] = b (+ imported update vectors); per front, in order:  S[struct] -= E S[own of the front];
     *             at the end  y = S[own] -> Z[n + row],  update vector of the root = S[boundary] -> Z[cl_ustore ...]
     *   backward  S[own] = y, S[boundary] = x of the ancestors; per front, in reverse:  S[own] = [F11^-1 | -G] [S[own]; S[struct]];
     *             at the end  x = S[own] -> Z[row] and the canonical state
     * so the update vectors of the fronts inside a cluster never exist in global memory.  Clusters of tier t import only
     * update vectors of cluster roots of lower tiers; all tiers run before the forward launches above and after the
     * backward launches.  ncluster may be 0.  Fronts are listed cluster by cluster in elimination order. */
    int32_t ntier;
    const int32_t* tier_ptr;   /* [ntier+1] cluster ranges                                          */
    int32_t ncluster;
    const int32_t* cl_fptr;    /* [ncluster+1] front ranges                                         */
    const int32_t* cl_ustore;  /* [ncluster] first Z row of the root's update vector, or -1         */
    const int64_t* cl_iptr;    /* [ncluster+1] ranges of imp_src / imp_dst                          */
    const int32_t* imp_src;    /* Z row of an imported update-vector entry                          */
    const int32_t* imp_dst;    /* solver row it is added to (a row resident in the cluster)         */
    int32_t nfront;
    const int32_t* fr_c0;      /* [nfront] first solver row of the front's own unknowns             */
    const int32_t* fr_w;       /* [nfront] own unknowns                                             */
    const int32_t* fr_m;       /* [nfront] boundary rows                                            */
    const int64_t* fr_sptr;    /* [nfront+1] ranges of fr_struct                                    */
    const int32_t* fr_struct;  /* solver rows of the boundary, increasing                           */
    const int64_t* fr_eptr;    /* [nfront] offset of E (m x w, row-major) in cl_vals                */
    const int64_t* fr_bptr;    /* [nfront] offset of [F11^-1 | -G] (w x (w+m), row-major) in cl_vals */
    const double* cl_vals;
} fcb_plan;

/* Everything that is constant over a run and shared by the whole ensemble.
 * Built on the host by flowcontrol_b200.problem.FlowProblem; replaces the
 * objects FlowSolver._setup/_prepare_systems create through dolfin
 * (src/flowcontrol/flowsolver.py:169-201, 665-701). */
typedef struct fcb_problem {
    /* mesh + P2 dof map (dolfin Mesh/FunctionSpace, flowsolver.py:233-250) */
    int32_t nT, nN, nV;
    const int32_t* cell_nodes; /* [nT*6] P2 node ids, local order v0 v1 v2 m12 m02 m01 */
    const double* Jinv;        /* [nT*4] row-major d(ref)/d(phys)                      */
    const double* detJ;        /* [nT] |det J|                                         */
    const double* node_xy;     /* [nN*2] P2 node coordinates, or NULL (only used to group cells and operator rows into compact patches) */
    /* unknown numbering */
    int32_t n_free;
    const int32_t* perm;    /* [n_free] solver row -> canonical dof in [0, 2nN+nV) */
    int32_t n_bc;
    const int32_t* bc_dofs; /* [n_bc] canonical Dirichlet dofs (DirichletBC, <example>flowsolver._make_bcs) */
    /* actuators (src/flowcontrol/actuator.py:184-313; flowsolver.py:278-309) */
    int32_t na;
    const double* bc_shape;     /* [na*n_bc] Dirichlet value per unit u_ctrl               */
    const double* ctrl_rhs[2];  /* per BDF order: [na*n_free] (force - lifting), solver rows */
    /* factorised LHS of BDF1 (index 0) and BDF2 (index 1) (nsforms.py:238-305) */
    fcb_plan plan[2];
    /* sensors as sparse rows over the canonical mixed vector (sensor.py:96-98,166-223) */
    int32_t ns;
    const int32_t* sensor_ptr; /* [ns+1] */
    const int32_t* sensor_idx;
    const double* sensor_val;
    /* scheme */
    double dt;
    int32_t nonlinear; /* ParamSolver.is_eq_nonlinear (nsforms.py:249,283-284) */
    /* ParamSolver.time_scheme (flowsolverparameters.py:192): 0 = "bdf" (BDF1 start-up, then BDF2), 1 = "cn"
     * (Crank-Nicolson, nsforms.py:191-236).  For "cn" both plans hold the same system, the per-step right-hand side is
     * E u_n - N(u_n) + ctrl_rhs[1] u_ctrl^{n+1} + ctrl_rhs_prev u_ctrl^n, with E = M/dt - (C + D + K/Re)/2 given as CSR. */
    int32_t scheme;
    const int32_t* cn_ptr;       /* [n_free+1] CSR of E, rows in solver order (pressure rows empty) */
    const int32_t* cn_idx;       /*            canonical velocity dof of each entry               */
    const double* cn_val;
    const double* ctrl_rhs_prev; /* [na*n_free] coefficient of the previous step's u_ctrl (averaged body force) */
} fcb_problem;

/* Controller bank: one discrete LTI controller per trajectory
 * (src/flowcontrol/controller.py:136-159):  v = Ky*y_meas ; u = Cd x + Dd v ;
 * x <- Ad x + Bd v ; u_ctrl = Fu*u.   Matrices are trajectory-innermost:
 * Ad[(i*nx+j)*B + b] etc.  Ky [ny*ns] and Fu [na*nu] are shared. */
typedef struct fcb_controllers {
    int32_t nx, ny, nu;
    const double* Ad; /* [nx*nx*B] */
    const double* Bd; /* [nx*ny*B] */
    const double* Cd; /* [nu*nx*B] */
    const double* Dd; /* [nu*ny*B] */
    const double* x0; /* [nx*B] or NULL (zeros) */
    const double* Ky; /* [ny*ns] */
    const double* Fu; /* [na*nu] */
} fcb_controllers;

/* Names of the phases timed by fcb_profile_step. */
enum { FCB_PHASE_RHS = 0, FCB_PHASE_FORWARD = 1, FCB_PHASE_BACKWARD = 2, FCB_PHASE_POST = 3,
       FCB_PHASE_SPMM = 4 /* Crank-Nicolson only */, FCB_PHASE_ELEMENT = 5, FCB_PHASE_MEASURE = 6, FCB_NPHASES = 7 };

/* Create an ensemble of B trajectories on CUDA device `device`.
 * Replaces FlowSolver._prepare_systems (flowsolver.py:665-701). */
int fcb_create(const fcb_problem* problem, int32_t B, int32_t device, fcb_handle* out);
int fcb_destroy(fcb_handle h);
const char* fcb_last_error(fcb_handle h);

/* Load the perturbation history (u_n, u_nn: [2nN*B]) and the BDF order (1 or 2)
 * of the next step.  Replaces FlowSolver.initialize_time_stepping
 * (flowsolver.py:464-549, 599-663).  Also evaluates y_meas and dE of u_n. */
int fcb_set_state(fcb_handle h, const double* u_n, const double* u_nn, const double* p_n, int32_t order);

int fcb_set_controllers(fcb_handle h, const fcb_controllers* c);

/* One time step of every trajectory.  u_ctrl [na*B] in; y_meas [ns*B], dE [B],
 * diverged [B] out (any may be NULL).  Replaces FlowSolver.step
 * (flowsolver.py:703-799): set_actuators_u_ctrl :724, assemble :728, solve :729,
 * split/_solver_diverged :730-731, history :746-751, make_measurement :761,
 * compute_perturbation_energy :775-779. */
int fcb_step(fcb_handle h, const double* u_ctrl, double* y_meas, double* dE, int32_t* diverged);

/* nsteps closed-loop steps without host round trips: controller -> step -> log.
 * series (may be NULL) receives [nsteps * ncol * B] with columns
 * (dE, u_ctrl_1..na, y_meas_1..ns) — the exporter's row minus time/runtime
 * (src/flowcontrol/exporter.py:191-224).  Requires fcb_set_controllers. */
int fcb_run_closed_loop(fcb_handle h, int32_t nsteps, double* series);

/* nsteps open-loop steps without host round trips: u_series [nsteps * na * B] holds the control inputs of every step
 * (host or device memory); series as above.  Replaces the user loop `for _ in range(num_steps): fs.step(u_ctrl(t))` of
 * the open-loop examples (src/examples/pinball/run_pinball_rotation_example.py:108-112,
 * src/examples/lidcavity/batch_run_lidcavity.py:203-216).  Both loops stream their series out in chunks of
 * FCB_SERIES_CHUNK steps (environment, default 1024): device memory for the series does not grow with nsteps. */
int fcb_run_open_loop(fcb_handle h, int32_t nsteps, const double* u_series, double* series);

/* Overwrite the controller states x [nx * B] (NULL = zeros): Controller.reset / restoring a checkpointed ensemble
 * (src/flowcontrol/controller.py:161-163).  fcb_set_state does NOT touch the controller states: like the reference's
 * restart (tests/integration/test_cylinder.py:95-112) the controller object lives on across a re-initialisation. */
int fcb_set_controller_state(fcb_handle h, const double* x);

/* Current fields in canonical numbering: up [ (2nN+nV) * B ].  which = 0: (u_, p_) of the
 * last step; 1: u_nn, only the 2nN velocity rows are written.  Replaces fs.fields.u_/p_/u_n/u_nn reads
 * (src/flowcontrol/flowfield.py:62-97). */
int fcb_get_fields(fcb_handle h, int32_t which, double* up);
int fcb_get_measurement(fcb_handle h, double* y_meas, double* dE, int32_t* diverged);
int fcb_get_controller_state(fcb_handle h, double* x);

/* Cost sums of the last fcb_run_closed_loop, accumulated on the device step by step: costs [3 * B] =
 * (sum_t dE, sum_t sum_k u_ctrl_k^2, dE of the last step).  Multiplied by Tnorm they are the reference's
 * compute_signal_cost(dE, 'integral' | 'terminal') and compute_control_cost (src/utils/optim.py:231-288); an
 * optimisation loop over controllers (fun_array, optim.py:48-66) reads 24 bytes per trajectory instead of the series. */
int fcb_get_costs(fcb_handle h, double* costs);

/* Mesh side of the matrix assembly: cells, their P2 nodes and geometry, a colouring in which cells of one colour share no
 * P2 node (flowcontrol_b200.mesh.TaylorHoodTables.element_colouring), and the position map of every element-matrix entry
 * in the CSR pattern of a scalar P2 operator (rows / columns = P2 nodes). */
typedef struct fcb_assembly {
    int32_t nT, nN;
    const int32_t* cell_nodes;   /* [nT*6] */
    const double* Jinv;          /* [nT*4] */
    const double* detJ;          /* [nT]   */
    int32_t ncolour;
    const int32_t* colour_ptr;   /* [ncolour+1] ranges of colour_cells */
    const int32_t* colour_cells; /* [nT] cells, colour by colour */
    int32_t nnz;                 /* entries of the scalar P2 pattern */
    const int32_t* pos;          /* [nT*36] entry (local row a, local column b) of cell e -> index into the pattern */
} fcb_assembly;

/* Assemble, for B velocity fields U [2nN * B] at once, the U-dependent blocks of the linearised operator on the scalar
 * P2 pattern: C [nnz * B] with C_ab = int (U.grad phi_b) phi_a, and (D != NULL) D [4 * nnz * B] = D^xx, D^xy, D^yx, D^yy
 * with D^ij_ab = int phi_b (d_j U_i) phi_a.  Replaces the dolfin assembly of the Jacobian / Picard operator in the
 * steady-state solver (src/flowcontrol/steadystate.py:95, 139-147 -> nsforms.py:137-183) and of the linearised
 * operator in OperatorGetter.get_A (src/flowcontrol/operatorgetter.py:25-83).  Stateless: needs no handle; host or
 * device pointers. */
int fcb_assemble_advection(const fcb_assembly* m, int32_t B, int32_t device, const double* U, double* C, double* D);

/* Symbolic side of the numeric multifrontal factorisation (flowcontrol_b200/devfactor.py: FrontMaps): the fronts of the
 * elimination tree (w own unknowns, m boundary rows; a front is the dense (w+m) x (w+m) matrix over [own | boundary]),
 * grouped in levels (children before parents), where the entries of the sparse matrix go, and where the boundary rows of
 * every child sit in its parent's front. */
typedef struct fcb_symbolic {
    int32_t nfront;
    const int32_t* w;            /* [nfront] */
    const int32_t* m;            /* [nfront] */
    int32_t nlevel;
    const int32_t* level_ptr;    /* [nlevel+1] ranges of level_fronts */
    const int32_t* level_fronts; /* [nfront] fronts, level by level */
    const int64_t* a_ptr;        /* [nfront+1] ranges of a_src / a_dst */
    const int32_t* a_src;        /* index into the CSR value array */
    const int32_t* a_dst;        /* row * (w+m) + column inside the front */
    const int64_t* c_ptr;        /* [nfront+1] ranges of c_front */
    const int32_t* c_front;      /* children */
    const int64_t* c_lptr;       /* [number of children + 1] ranges of c_loc */
    const int32_t* c_loc;        /* position of each boundary row of the child in the parent's front */
} fcb_symbolic;

/* Numeric multifrontal factorisation on the GPU: for every front, F11^-1 (w x w, Gauss-Jordan with partial pivoting),
 * E = F21 F11^-1 (m x w), G = F11^-1 F12 (w x m), all row-major and concatenated front by front, plus growth[nfront] =
 * max|F11^-1| max|F11| (singularity check).  avals [nnz]: values of the permuted matrix in CSR order.  Replaces the numeric
 * phase of MUMPS behind dolfin.LUSolver.set_operator / dolfin.solve (src/flowcontrol/flowsolver.py:694-697,
 * steadystate.py:95, 142-144).  Stateless; host or device pointers. */
int fcb_factorize(const fcb_symbolic* s, int32_t device, const double* avals, int64_t nnz, double* E, double* Finv, double* G,
                  double* growth);

/* Runs one step with CUDA events between phases; ms[FCB_NPHASES] receives device times,
 * launches[FCB_NPHASES] (may be NULL) the kernel launches per phase. */
int fcb_profile_step(fcb_handle h, const double* u_ctrl, float* ms, int32_t* launches);

/* Kernel launches issued by this handle since creation (graph replays count their nodes). */
int64_t fcb_launch_count(fcb_handle h);
/* Raw cudaStream_t of the handle (for CUDA-event timing by the caller). */
void* fcb_stream(fcb_handle h);
int fcb_synchronize(fcb_handle h);
/* Library/version string, e.g. "fcb200 0.1 sm_100a". */
const char* fcb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FCB200_H */
