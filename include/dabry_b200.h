/* dabry_b200 - C ABI of the B200-native extremal-field solver (libdabry_b200.so).
 *
 * The reference (dabry v1.1) is pure Python and has no FFI; the seam this library plugs into is the solver class
 * interface `SolverEF*.setup()/step()/solve()` (dabry/solver_ef.py:281-857).  Each entry point below names the
 * reference code it replaces.  Conventions: plain C, every pointer is a HOST pointer owned by the caller unless
 * stated, every function returns 0 on success and a negative dabry_status on error, `dabry_last_error` gives the
 * message.  A handle is bound to one CUDA device and one stream and is not thread-safe.  Different handles may be
 * driven from different host threads: the calls of all handles on ONE device are serialised inside the library (they
 * share the constant memory the problem description lives in), handles on different devices run concurrently.  There
 * is no host execution path: dabry_create fails with DABRY_ERR_NO_DEVICE when no CUDA device is usable.
 */
#ifndef DABRY_B200_H
#define DABRY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DABRY_ABI_VERSION 2
#define DABRY_MAX_OBSTACLES 8
#define DABRY_MAX_LEAVES 8
#define DABRY_MAX_TARGET_EVENTS 4

typedef struct dabry_solver dabry_solver;
typedef struct dabry_comm dabry_comm; /* a communicator over the GPUs of one job (NCCL), see "in-library collectives" */
#define DABRY_COMM_ID_BYTES 128

enum dabry_status {
    DABRY_OK = 0,
    DABRY_ERR_INVALID = -1,    /* bad argument / call order */
    DABRY_ERR_NO_DEVICE = -2,  /* no usable CUDA device */
    DABRY_ERR_CUDA = -3,       /* a CUDA runtime call or kernel failed */
    DABRY_ERR_CAPACITY = -4,   /* more sites than config.max_sites (when growth is disabled) or than memory */
    DABRY_ERR_UNSUPPORTED = -5
};

enum dabry_coords { DABRY_CARTESIAN = 0, DABRY_GCS = 1 };                 /* misc.py:142-154 */
enum dabry_integrator { DABRY_DOPRI5_SCIPY = 0, DABRY_RK4_FIXED = 1 };     /* scipy RK45 restated | classic RK4 */
enum dabry_dtype { DABRY_F64 = 0, DABRY_F32 = 1 };
enum dabry_variant { DABRY_SIMPLE = 0, DABRY_RESAMPLING = 1, DABRY_TRIMMING = 2 }; /* solver_ef.py:406, 447, 762 */

/* Closure / neutering flags are the reference's enum values (solver_ef.py:42-53); -1 = None. */

/* SolverEF.__init__ + SolverEFResampling/Trimming.__init__ constants (solver_ef.py:284-339, 449-460, 764-776). */
typedef struct dabry_config {
    int32_t abi_version;          /* DABRY_ABI_VERSION */
    int32_t coords;               /* dabry_coords: picks dyn_augsys_cartesian / _gcs (solver_ef.py:339) */
    int32_t n_time;               /* len(solver.times) */
    int32_t max_depth;
    int32_t mode_origin;          /* solver_ef.py:449 */
    int32_t integrator;           /* dabry_integrator */
    int32_t dtype;                /* dabry_dtype */
    int32_t device;               /* CUDA device ordinal */
    int32_t n_index_per_subframe; /* trimming (solver_ef.py:765) */
    int32_t trimming_band_width;  /* trimming (solver_ef.py:766) */
    int32_t cost_map_nx, cost_map_ny; /* 100 x 100 in the reference (solver_ef.py:820) */
    int64_t n_roots;              /* n_costate_sectors: size of the whole ring of root extremals */
    int64_t root_offset;          /* multi-GPU: this handle owns roots [root_offset, root_offset + root_count) */
    int64_t root_count;           /* == n_roots on a single GPU */
    int64_t max_sites;            /* initial capacity in sites (0: 2 * root_count + 1024); storage grows on demand */
    int64_t shard_block;          /* 0: contiguous shard.  > 0: block-cyclic - this handle owns blocks of shard_block roots,
                                     every shard_world-th block of the ring, the first one starting at root_offset */
    int64_t shard_world;
    double t_init;                /* cost = t - t_init (solver_ef.py:501) */
    double max_step;              /* solver.max_int_step (solver_ef.py:335-337) */
    double rtol, atol;            /* scipy defaults 1e-3, 1e-6 (rk.py:86) */
    double srf;                   /* pb.srf_max */
    double x_init[2], x_target[2];
    double target_radius;
    double max_dist;              /* resampling threshold (solver_ef.py:451) */
    double bl[2], tr[2];          /* problem box: cost-map extent (solver_ef.py:825-829) */
    double ff_max_norm;           /* max |w| over the grid nodes (solver_ef.py:459); 0 disables trim_distance (594-600);
                                     < 0: use the value dabry_set_flow_grid computes on the device */
    const double* times;          /* [n_time] solver.times = linspace(t_init, t_init + T, n_time) (solver_ef.py:313) */
} dabry_config;

/* WrapperFF parameters (flowfield.py:522-544); identity = {0, 1, 1, {0, 0}, 1, 1}. */
typedef struct dabry_wrapper {
    double time_origin, scale_time, scale_length, bl[2], scaler_speed, scaler_dspeed;
} dabry_wrapper;

/* One analytic leaf of a flattened FlowField expression tree (flowfield.py:66-92): value = sum_i coeff_i leaf_i. */
enum dabry_leaf_kind {
    DABRY_LEAF_UNIFORM = 0,      /* p: value(2)                                      flowfield.py:601-620 */
    DABRY_LEAF_STATE_LINEAR = 1, /* p: gradient(4, row major), origin(2), value_origin(2)   :730-752, :783-804 */
    DABRY_LEAF_GYRE = 2,         /* p: center(2), kx, ky, ampl                       :807-833 */
    DABRY_LEAF_VORTEX = 3,       /* p: center(2), circulation                        :623-649 */
    DABRY_LEAF_RANKINE = 4,      /* p: center(2), circulation, radius | table rows (cx, cy, gamma, radius)  :652-727 */
    DABRY_LEAF_CHERTOVSKIH = 5,  /*                                                  :1089-1099 */
    DABRY_LEAF_TRAP = 6,         /* p[0]: rel_wid; table rows (ff_val, cx, cy, radius)      :1035-1086 */
    DABRY_LEAF_RADIAL_GAUSS = 7, /* p: center(2), radius, sdev, v_max                :860-903 */
    DABRY_LEAF_TWO_SECTORS = 8,  /* p: v_w1, v_w2, x_switch (TwoSectorsFF, TSEqualFF)  :565-598 */
    DABRY_LEAF_LINEAR_T = 9,     /* table of 3 rows: gradient_diff_time(4), gradient_init(4), origin(2) value_origin(2)  :755-780 */
    DABRY_LEAF_RADIAL_GAUSS_T = 10, /* table: nt rows (cx, cy, radius, sdev), then nt rows (v_max, 0, 0, 0)   :906-973 */
    DABRY_LEAF_BAND_GAUSS = 11,  /* p: origin(2), unit vect(2), ampl, sdev; Jacobian by forward differences as written  :976-999 */
    DABRY_LEAF_BAND = 12,        /* p: origin(2), unit vect(2), w_value(2), width; central differences as written  :1002-1032 */
    DABRY_LEAF_GYRE_MSEAS = 13   /* p: A, eps, T; central differences as written     :1102-1134 */
};

typedef struct dabry_leaf {
    int32_t kind;
    int32_t nt;        /* > 0: time table of nt rows x 4 doubles (FlowField._index, flowfield.py:152-175) */
    int32_t table_off; /* offset in doubles into the table buffer */
    int32_t pad_;
    double coeff;
    double t_start, t_end;
    double p[8];
} dabry_leaf;

enum dabry_obstacle_kind { DABRY_OBS_CIRCLE = 0, DABRY_OBS_FRAME = 1, DABRY_OBS_GREAT_CIRCLE = 2 }; /* obstacle.py:84-172 */

typedef struct dabry_obstacle {
    int32_t kind;
    int32_t pad_;
    double scale_length, bl[2]; /* WrapperObs (obstacle.py:66-81); identity = 1, {0, 0} */
    double p[8];                /* circle: center(2), radius**2 | frame: center(2), scaler diagonal(2) |
                                   great circle: dir_vect(3), zone z1(2), z2(2), 1 if the zone is set else 0 */
} dabry_obstacle;

enum dabry_penalty_kind { DABRY_PEN_NONE = 0, DABRY_PEN_CIRCLE = 1, DABRY_PEN_GRID = 2 }; /* penalty.py:72-134 */

typedef struct dabry_penalty {
    int32_t kind;
    int32_t pad_;
    double time_origin, scale_time, scale_length, bl[2], scaler; /* WrapperPen (penalty.py:47-69) */
    double dx;   /* finite-difference step (penalty.py:33) */
    double p[4]; /* circle: center(2), radius**2, amplitude */
} dabry_penalty;

/* Per-site record, the fields of class Site (solver_ef.py:56-164) that are not trajectories. */
typedef struct dabry_site_record {
    int64_t dyadic;          /* SiteManager index (solver_ef.py:197-202): names are derived from it on the host */
    int32_t index_t_init;
    int32_t n_samples;       /* len(site.traj) */
    int32_t closure;         /* ClosureReason value or -1 */
    int32_t neuter;          /* NeuteringReason value or -1 */
    int32_t index_neutered;
    int32_t obstacle;        /* index into the event table (1 = first obstacle), -1 = free */
    int32_t index_t_obs;
    int32_t obs_trigo;
    int32_t index_t_check_next;
    int32_t depth;
    int32_t parent_prev, parent_next; /* slots of the parents, -1 for roots */
    int32_t n_target_events;
    int32_t ivp_status;
    int32_t flags;           /* bit 0: ghost copy of a neighbouring rank's root (multi-GPU), not a site of this rank */
    int32_t pad_;
    double t_enter_obs;      /* NaN = None */
    double cost0;            /* cost of the first sample */
    double t_target_events[DABRY_MAX_TARGET_EVENTS];
} dabry_site_record;

typedef struct dabry_stats {
    uint64_t rk_steps;       /* RK step attempts (accepted + rejected), summed over extremals: the metric's unit */
    uint64_t rhs_evals;
    uint64_t ivp_calls;
    uint64_t kernel_launches;
    int64_t n_sites;
    int64_t capacity;
    int32_t rounds;
    int32_t pad_;
    double ms_integrate, ms_resample, ms_trim, ms_arrival, ms_total;   /* device time by phase (CUDA events); ms_integrate = free-arc kernel */
    double ms_upload;        /* flow grid repack + derivative kernels */
    double ms_obstacle;      /* obstacle entry + boundary-following kernel (captured extremals only) */
    double ms_comm;          /* in-library collectives of dabry_solve_sharded (events on the solver's stream) */
    uint64_t rk_steps_free;  /* the share of rk_steps taken inside the free-arc kernel (what ms_integrate times) */
} dabry_stats;

/* ---- lifetime ------------------------------------------------------------------------------------------ */
int dabry_abi_version(void);
/* SolverEF.__init__ (solver_ef.py:284-339): the host computes the constants, the library copies them. */
int dabry_create(const dabry_config* config, dabry_solver** out);
void dabry_destroy(dabry_solver* h);
const char* dabry_last_error(const dabry_solver* h); /* h may be NULL: error of the last failed dabry_create */
const char* dabry_backend_name(void);                /* "cuda-sm100a" for the product library */

/* ---- host staging --------------------------------------------------------------------------------------- */
/* Page-lock (pin) a caller-owned host array in place / release it.  dabry_set_flow_grid and dabry_setup accept any
 * host pointer; from a pinned array the H2D copy is one DMA at the PCIe rate, from pageable memory the driver stages
 * it through its own bounce buffers (measured 4x slower for the 0.4 GB of a 1440 x 721 x 24 grid).  No reference
 * counterpart (numpy arrays are the reference's only storage, flowfield.py:183-214). */
int dabry_host_register(void* p, size_t bytes);
int dabry_host_unregister(void* p);
/* The library parks the large device blocks of a destroyed solver in a per-device cache so that the next solver of the
 * same shape starts without allocating (one handle per solve() is the reference's usage pattern).  This gives them back
 * to the driver.  DABRY_B200_CACHE_MB in the environment bounds the cache (0 disables it). */
int dabry_release_cached_memory(int32_t device);

/* ---- problem ------------------------------------------------------------------------------------------- */
/* DiscreteFF (flowfield.py:178-214) seen through WrapperFF (518-544).  values: C-order (nt, nx, ny, 2), or
 * (nx, ny, 2) when unsteady == 0 (then nt must be 1).  grad: (..., 2, 2) or NULL -> compute_derivatives
 * (flowfield.py:474-504) runs on the device.  lo/hi: bounds[:, 0] / bounds[:, 1] for (t, x, y) (t ignored when
 * steady).  The library repacks into its own HBM layout; the caller's arrays are not retained - with one exception:
 * a PAGE-LOCKED `values` array (dabry_host_register) of 64 MB or more of an unsteady field without `grad` is uploaded in the
 * background in groups of time frames, and the first solve packs and uses them as they arrive (its free arcs are
 * parked at the time the frames on the device run out).  Such an array must stay alive and unchanged until the next
 * call that needs the whole grid returns (dabry_solve / dabry_step*, dabry_eval_*, dabry_get_flow_max_norm).
 * DABRY_B200_STREAM_UPLOAD=0 in the environment restores the one-shot upload. */
int dabry_set_flow_grid(dabry_solver* h, const double* values, const double* grad, int32_t nt, int32_t nx,
                        int32_t ny, int32_t unsteady, const double lo[3], const double hi[3],
                        const dabry_wrapper* wrapper);
/* Analytic FlowField tree (flowfield.py:39-150, 565-1134), flattened by the host. */
int dabry_set_flow_program(dabry_solver* h, const dabry_leaf* leaves, int32_t n_leaves, const double* tables,
                           int64_t n_table_doubles, const dabry_wrapper* wrapper);
/* pb.obstacles in order, the auto frame last (problem.py:100-104; solver_ef.py:322-334). */
int dabry_set_obstacles(dabry_solver* h, const dabry_obstacle* obstacles, int32_t n);
int dabry_set_penalty(dabry_solver* h, const dabry_penalty* penalty);
/* DiscretePenalty (penalty.py:96-134): data [nt][nx][ny] sampled on the (not necessarily uniform) axes ts / xs / ys,
 * interpolated like scipy's RegularGridInterpolator(method='linear', bounds_error=False, fill_value=0) and
 * differentiated by central differences with penalty->dx like Penalty.d_value (penalty.py:39-44); `penalty` carries
 * the WrapperPen parameters and dx (its kind and p[] are ignored).  The arrays are copied. */
int dabry_set_penalty_grid(dabry_solver* h, const dabry_penalty* penalty, const double* data, const double* ts,
                           const double* xs, const double* ys, int32_t nt, int32_t nx, int32_t ny);

/* ---- solve --------------------------------------------------------------------------------------------- */
/* SolverEFResampling.setup (solver_ef.py:462-470) / SolverEFTrimming.setup (778-782) / the fan of
 * SolverEFSimple.step (419-424).  costates: [root_count][2] initial costates of this handle's roots, or NULL ->
 * 1000 (cos, sin)(theta_i), theta_i = linspace(0, 2 pi, n_roots, endpoint=False)[root_offset + i] on the device. */
int dabry_setup(dabry_solver* h, int32_t variant, const double* costates);
/* One SolverEF*.step() (solver_ef.py:419-437, 582-592, 802-817). */
int dabry_step(dabry_solver* h);
/* The two halves of a Resampling/Simple step, for callers that exchange the ring boundary in between (multi-GPU):
 * integrate + connect_to_parents (solver_ef.py:583-587), then compute_new_sites + trim_distance (589-591). */
int dabry_step_integrate(dabry_solver* h);
int dabry_step_resample(dabry_solver* h);
/* The three parts of a Trimming step (solver_ef.py:802-841), for callers that exchange the ring boundary after the
 * first and all_reduce(MIN) the cost map (dabry_get_cost_map / dabry_set_cost_map) after the second:
 * integrate sites_valid[i] (803-807) | compute_new_sites fixed point (808-815) + local cost_map_triangle (821-828) |
 * cost-map lookup and SUBOPTIMAL closure (829-841). */
int dabry_step_trim_integrate(dabry_solver* h);
int dabry_step_trim_refine(dabry_solver* h);
int dabry_step_trim_close(dabry_solver* h);
int dabry_set_cost_map(dabry_solver* h, const double* map);
/* check_solutions (solver_ef.py:667-685). */
int dabry_finish(dabry_solver* h);
/* setup + all steps + finish: SolverEFResampling.solve (602-608) / SolverEFTrimming.solve (843-857). */
int dabry_solve(dabry_solver* h, int32_t variant, const double* costates);

/* ---- results ------------------------------------------------------------------------------------------- */
int dabry_num_sites(dabry_solver* h, int64_t* n);
int dabry_get_sites(dabry_solver* h, dabry_site_record* out, int64_t first, int64_t count);
/* Trajectories of sites [first, first + count), each padded to n_time slots; slot k = time-grid index k, valid for
 * index_t_init <= k < index_t_init + n_samples.  Any output pointer may be NULL.
 * times [count][n_time], states/costates/controls [count][n_time][2], cost [count][n_time]. */
int dabry_get_trajs(dabry_solver* h, int64_t first, int64_t count, double* times, double* states,
                    double* costates, double* controls, double* cost);
/* next_nb [count][n_time] slots (-1 = None) (solver_ef.py:79). */
int dabry_get_next_nb(dabry_solver* h, int64_t first, int64_t count, int32_t* next_nb);
/* Per-site arrival: grid index of the cheapest in-target sample (-1: never inside) and its cost. */
int dabry_get_arrivals(dabry_solver* h, int64_t first, int64_t count, int32_t* index, double* cost);
/* solution_sites (solver_ef.py:676-685): ascending slots of the sites that arrived with a cost not above `bound`
 * (the caller passes 1.15 x the optimum), compacted on the device.  *count is always set; slots (capacity cap, may be
 * NULL to query the count) is filled when count <= cap. */
int dabry_get_solution_sites(dabry_solver* h, double bound, int32_t* slots, int64_t cap, int64_t* count);
/* solution_site (slot, -1 if none) and its cost; ties -> lowest slot. */
int dabry_get_solution(dabry_solver* h, int64_t* site, double* cost);
/* The last trimming cost map [cost_map_nx][cost_map_ny] (solver_ef.py:828), or cost_map_triangle over all sites
 * and all times when full != 0 (solver_ef.py:687-688). */
int dabry_get_cost_map(dabry_solver* h, int32_t full, double* out);
int dabry_get_stats(dabry_solver* h, dabry_stats* out);
/* np.max(np.linalg.norm(ff.values, axis=-1)) of the wrapped grid (solver_ef.py:459), computed while repacking. */
int dabry_get_flow_max_norm(dabry_solver* h, double* out);

/* ---- multi-GPU ring boundary ---------------------------------------------------------------- */
/* The last root of every local block needs the next root of the ring, owned by the next rank, as its neighbour
 * (solver_ef.py:468-469).  A packet = one root trajectory, dabry_boundary_doubles doubles ([n_time][7] samples +
 * header); a handle has dabry_boundary_count packets (= its number of blocks; 1 for a contiguous shard).
 * export: the FIRST root of each local block, in block order.  import: packets installed as the ghost sites of the
 * local blocks, in block order (the caller routes rank r+1's packets to rank r, see dabry_b200/distributed.py). */
int dabry_boundary_doubles(dabry_solver* h, int64_t* n);
int dabry_boundary_count(dabry_solver* h, int64_t* n);
int dabry_export_boundary(dabry_solver* h, double* packed);
int dabry_import_boundary(dabry_solver* h, const double* packed);
/* The same with DEVICE buffers of the caller on the handle's device - the tensors NCCL sends from / receives into - so
 * that the exchange never touches the host: packets are written / read in place, the cost map (solver_ef.py:825-828) is
 * copied device to device.  The caller synchronises its own stream before the call; the call returns with the
 * handle's stream drained, so the buffer can go straight to the collective. */
int dabry_export_boundary_device(dabry_solver* h, double* d_packed);
int dabry_import_boundary_device(dabry_solver* h, const double* d_packed);
int dabry_get_cost_map_device(dabry_solver* h, int32_t full, double* d_out);
int dabry_set_cost_map_device(dabry_solver* h, const double* d_map);

/* ---- in-library collectives ------------------------------------------------------------------ */
/* One process per GPU.  The ring of roots is sharded by theta_0 through the dabry_config fields root_offset, root_count,
 * shard_block and shard_world, the flow grid is replicated, and the library itself carries the three exchanges of the
 * reference's loops over NCCL, enqueued on the solver's stream with no host round trip:
 *   - ring boundary (solver_ef.py:468-469): ncclSend / ncclRecv of one packet per local block inside one
 *     ncclGroupStart / ncclGroupEnd, after the roots are integrated (Trimming: every sub-frame);
 *   - trimming cost map (solver_ef.py:825-828): ncclAllReduce(ncclMin) in place on the device map, every sub-frame;
 *   - check_solutions (solver_ef.py:667-685): ncclAllGather of (arrival cost, step count) per rank at the end.
 * NCCL is bound at run time (the process's own copy - e.g. PyTorch's - else libnccl.so.2); without it
 * dabry_comm_unique_id / dabry_comm_create return DABRY_ERR_UNSUPPORTED and dabry_comm_last_error says why.
 * Rank 0 calls dabry_comm_unique_id and hands the DABRY_COMM_ID_BYTES to the other ranks by any means (the Python
 * host mirror broadcasts them through torch.distributed); every rank then calls dabry_comm_create, which is collective.
 * A communicator outlives the solvers attached to it (one solver per solve() is the reference's pattern). */
int dabry_comm_unique_id(void* id);
int dabry_comm_create(int32_t device, int32_t rank, int32_t world, const void* id, dabry_comm** out);
void dabry_comm_destroy(dabry_comm* comm);
const char* dabry_comm_last_error(void);
int dabry_attach_comm(dabry_solver* h, dabry_comm* comm); /* NULL detaches; the handle does not own it */
/* dabry_set_flow_grid for ranks that all hold the same `values` (grad is computed on the device): each rank uploads
 * 1 / world of the samples over PCIe and the rest arrives from its peers over NVLink (ncclAllGather in place).
 * Collective over the attached communicator; without one it is dabry_set_flow_grid. */
int dabry_set_flow_grid_shared(dabry_solver* h, const double* values, int32_t nt, int32_t nx, int32_t ny,
                               int32_t unsteady, const double lo[3], const double hi[3], const dabry_wrapper* wrapper);
/* dabry_solve over all ranks of the attached communicator (collective); without one it is dabry_solve. */
int dabry_solve_sharded(dabry_solver* h, int32_t variant, const double* costates);
/* Global optimum after dabry_solve_sharded: arrival cost (NaN: none), the rank that holds the optimal extremal (lowest
 * rank among equals, -1: none) and the RK step attempts of all ranks. */
int dabry_get_global_solution(dabry_solver* h, double* cost, int32_t* owner_rank, uint64_t* total_rk_steps);

/* ---- time horizon --------------------------------------------------------------------- */
/* NavigationProblem.auto_time_upper_bound (problem.py:193-239, feedback.py:58-130): the two feedback flights from x_init
 * - ground velocity aimed at the target (GSTargetFB) and heading aimed at the target (HTargetFB, a unit control as
 * written) - integrated on the device by the same RK45 / event root finder as the extremals (scipy defaults, no step
 * limit, terminal target event of either direction) over [t_start, t_end] = [ff.t_start, ff.t_start + 10 L / srf].
 * out[0] = time_orthodromic, out[1] = time_htarget; +inf when the target disc is not reached.  Uses the handle's flow
 * field, coords, srf, x_target, target_radius, rtol and atol. */
int dabry_time_upper_bound(dabry_solver* h, const double x_init[2], double t_start, double t_end, double out[2]);

/* ---- evaluation hooks used by the parity tests ---------------------------------------------------------- */
/* SolverEF.dyn_augsys(t, y) (solver_ef.py:355-359) for n points: t [n], y [n][4] -> dy [n][4]. */
int dabry_eval_rhs(dabry_solver* h, int64_t n, const double* t, const double* y, double* dy);
/* ff.value / ff.d_value (flowfield.py:538-544): w [n][2], jac [n][2][2]. */
int dabry_eval_flow(dabry_solver* h, int64_t n, const double* t, const double* x, double* w, double* jac);

#ifdef __cplusplus
}
#endif
#endif /* DABRY_B200_H */
