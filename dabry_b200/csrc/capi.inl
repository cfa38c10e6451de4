// extern "C" entry points of include/dabry_b200.h over Engine<DABRY_BACKEND>.
// Included once by dabry_b200.cu (DABRY_BACKEND = CudaBackend: the product library) and once by
// tests/hostsim/hostsim.cpp (DABRY_BACKEND = SerialBackend: g++-built test double, never shipped in the package).
#ifndef DABRY_BACKEND
#error "define DABRY_BACKEND before including capi.inl"
#endif

struct dabry_solver {
    dabry::Engine<DABRY_BACKEND>* eng;
};

static thread_local std::string g_create_error;

struct dabry_comm {
    DABRY_BACKEND::Comm* comm;
};

static thread_local std::string g_comm_error;

// every entry point first takes the call guard of the handle's backend: it makes the handle's device current (a process
// may hold handles on several GPUs) and serialises the calls of all handles on that device, which share the constant
// memory the problem description lives in (CudaBackend::CallGuard)
#define DABRY_GUARD(h)                 \
    if (!(h) || !(h)->eng) return DABRY_ERR_INVALID; \
    dabry::Engine<DABRY_BACKEND>& E = *(h)->eng;     \
    DABRY_BACKEND::CallGuard call_guard_(E.be)

extern "C" {

int dabry_abi_version(void) { return DABRY_ABI_VERSION; }

const char* dabry_backend_name(void) { return DABRY_BACKEND::name(); }

int dabry_host_register(void* p, size_t bytes) {
    if (!p || bytes == 0) return DABRY_ERR_INVALID;
    return DABRY_BACKEND::host_register(p, bytes) ? DABRY_OK : DABRY_ERR_CUDA;
}

int dabry_host_unregister(void* p) {
    if (!p) return DABRY_ERR_INVALID;
    return DABRY_BACKEND::host_unregister(p) ? DABRY_OK : DABRY_ERR_CUDA;
}

int dabry_release_cached_memory(int32_t device) {
    return DABRY_BACKEND::release_cached(device) ? DABRY_OK : DABRY_ERR_INVALID;
}

int dabry_create(const dabry_config* c, dabry_solver** out) {
    if (!c || !out) return DABRY_ERR_INVALID;
    *out = nullptr;
    if (c->abi_version != DABRY_ABI_VERSION) {
        g_create_error = "dabry_config.abi_version mismatch";
        return DABRY_ERR_INVALID;
    }
    if (c->n_time < 2 || !c->times || c->n_roots < 1 || c->max_depth < 0 || c->max_depth > 40 ||
        c->cost_map_nx < 2 || c->cost_map_ny < 2 || c->n_index_per_subframe < 1) {
        g_create_error = "dabry_config: need n_time >= 2, times, n_roots >= 1, 0 <= max_depth <= 40, cost map >= 2x2";
        return DABRY_ERR_INVALID;
    }
    if ((c->dtype != DABRY_F64 && c->dtype != DABRY_F32) ||
        (c->integrator != DABRY_DOPRI5_SCIPY && c->integrator != DABRY_RK4_FIXED)) {
        g_create_error = "unknown dtype / integrator";
        return DABRY_ERR_UNSUPPORTED;
    }
    // DABRY_B200_DEBUG_TIMING=1: host wall time of the two halves of dabry_create on stderr (diagnostic)
    const bool timing = getenv("DABRY_B200_DEBUG_TIMING") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    // dyadic site indices are root * 2^max_depth (solver_ef.py:197-202) in int64
    if (c->max_depth >= 62 || c->n_roots > (((int64_t) 1 << 62) >> c->max_depth)) {
        g_create_error = "dabry_config: n_roots * 2^max_depth does not fit the 64-bit dyadic site index";
        return DABRY_ERR_INVALID;
    }
    dabry::Engine<DABRY_BACKEND>* e = new dabry::Engine<DABRY_BACKEND>(*c);
    if (!e->be.open(c->device)) {
        g_create_error = e->be.error();
        delete e;
        return DABRY_ERR_NO_DEVICE;
    }
    const auto t1 = std::chrono::steady_clock::now();
    const int rc = e->init();
    if (timing)
        fprintf(stderr, "dabry_create: open %.2f ms, init %.2f ms\n",
                std::chrono::duration<double, std::milli>(t1 - t0).count(),
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t1).count());
    if (rc != DABRY_OK) {
        g_create_error = e->err;
        delete e;
        return rc;
    }
    *out = new dabry_solver{e};
    return DABRY_OK;
}

void dabry_destroy(dabry_solver* h) {
    if (!h) return;
    const bool timing = getenv("DABRY_B200_DEBUG_TIMING") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    if (h->eng) DABRY_BACKEND::destroy_locked(h->eng->be, [&] { delete h->eng; });
    delete h;
    if (timing)
        fprintf(stderr, "dabry_destroy: %.2f ms\n",
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
}

const char* dabry_last_error(const dabry_solver* h) {
    if (!h || !h->eng) return g_create_error.c_str();
    return h->eng->err.c_str();
}

int dabry_set_flow_grid(dabry_solver* h, const double* values, const double* grad, int32_t nt, int32_t nx, int32_t ny,
                        int32_t unsteady, const double lo[3], const double hi[3], const dabry_wrapper* w) {
    DABRY_GUARD(h);
    if (!lo || !hi) return E.fail(DABRY_ERR_INVALID, "flow grid: bounds missing");
    return E.set_flow_grid(values, grad, nt, nx, ny, unsteady, lo, hi, w);
}

int dabry_set_flow_grid_shared(dabry_solver* h, const double* values, int32_t nt, int32_t nx, int32_t ny, int32_t unsteady,
                               const double lo[3], const double hi[3], const dabry_wrapper* w) {
    DABRY_GUARD(h);
    if (!lo || !hi) return E.fail(DABRY_ERR_INVALID, "flow grid: bounds missing");
    return E.set_flow_grid(values, nullptr, nt, nx, ny, unsteady, lo, hi, w, true);
}

int dabry_set_flow_program(dabry_solver* h, const dabry_leaf* leaves, int32_t n, const double* tables, int64_t n_tab,
                           const dabry_wrapper* w) {
    DABRY_GUARD(h);
    if (!leaves) return E.fail(DABRY_ERR_INVALID, "flow program: leaves missing");
    return E.set_flow_program(leaves, n, tables, n_tab, w);
}

int dabry_set_obstacles(dabry_solver* h, const dabry_obstacle* obs, int32_t n) {
    DABRY_GUARD(h);
    return E.set_obstacles(obs, n);
}

int dabry_set_penalty(dabry_solver* h, const dabry_penalty* pen) {
    DABRY_GUARD(h);
    if (!pen) return E.fail(DABRY_ERR_INVALID, "penalty missing");
    return E.set_penalty(pen);
}

int dabry_set_penalty_grid(dabry_solver* h, const dabry_penalty* pen, const double* data, const double* ts,
                           const double* xs, const double* ys, int32_t nt, int32_t nx, int32_t ny) {
    DABRY_GUARD(h);
    if (!pen) return E.fail(DABRY_ERR_INVALID, "penalty missing");
    return E.set_penalty_grid(pen, data, ts, xs, ys, nt, nx, ny);
}

int dabry_setup(dabry_solver* h, int32_t variant, const double* costates) {
    DABRY_GUARD(h);
    return E.setup(variant, costates);
}

int dabry_step(dabry_solver* h) {
    DABRY_GUARD(h);
    return E.step();
}

int dabry_step_integrate(dabry_solver* h) {
    DABRY_GUARD(h);
    return E.step_integrate();
}

int dabry_step_resample(dabry_solver* h) {
    DABRY_GUARD(h);
    return E.step_resample();
}

int dabry_step_trim_integrate(dabry_solver* h) {
    DABRY_GUARD(h);
    return E.step_trim_integrate();
}

int dabry_step_trim_refine(dabry_solver* h) {
    DABRY_GUARD(h);
    return E.step_trim_refine();
}

int dabry_step_trim_close(dabry_solver* h) {
    DABRY_GUARD(h);
    return E.step_trim_close();
}

int dabry_set_cost_map(dabry_solver* h, const double* map) {
    DABRY_GUARD(h);
    if (!map) return DABRY_ERR_INVALID;
    return E.set_cost_map(map);
}

int dabry_finish(dabry_solver* h) {
    DABRY_GUARD(h);
    return E.finish();
}

int dabry_solve(dabry_solver* h, int32_t variant, const double* costates) {
    DABRY_GUARD(h);
    return E.solve(variant, costates);
}

int dabry_num_sites(dabry_solver* h, int64_t* n) {
    DABRY_GUARD(h);
    if (!n) return DABRY_ERR_INVALID;
    *n = E.num_sites();
    return DABRY_OK;
}

int dabry_get_sites(dabry_solver* h, dabry_site_record* out, int64_t first, int64_t count) {
    DABRY_GUARD(h);
    return E.get_sites(out, first, count);
}

int dabry_get_trajs(dabry_solver* h, int64_t first, int64_t count, double* times, double* states, double* costates,
                    double* controls, double* cost) {
    DABRY_GUARD(h);
    return E.get_trajs(first, count, times, states, costates, controls, cost);
}

int dabry_get_next_nb(dabry_solver* h, int64_t first, int64_t count, int32_t* next_nb) {
    DABRY_GUARD(h);
    return E.get_next_nb(first, count, next_nb);
}

int dabry_get_arrivals(dabry_solver* h, int64_t first, int64_t count, int32_t* index, double* cost) {
    DABRY_GUARD(h);
    return E.get_arrivals(first, count, index, cost);
}

int dabry_get_solution(dabry_solver* h, int64_t* site, double* cost) {
    DABRY_GUARD(h);
    if (site) *site = E.solution();
    if (cost) *cost = E.solution_cost_value();
    return DABRY_OK;
}

int dabry_get_solution_sites(dabry_solver* h, double bound, int32_t* slots, int64_t cap, int64_t* count) {
    DABRY_GUARD(h);
    if (!count || cap < 0) return DABRY_ERR_INVALID;
    return E.get_solution_sites(bound, slots, cap, count);
}

int dabry_get_cost_map(dabry_solver* h, int32_t full, double* out) {
    DABRY_GUARD(h);
    if (!out) return DABRY_ERR_INVALID;
    return E.get_cost_map(full, out);
}

int dabry_get_stats(dabry_solver* h, dabry_stats* out) {
    DABRY_GUARD(h);
    if (!out) return DABRY_ERR_INVALID;
    return E.get_stats(out);
}

int dabry_get_flow_max_norm(dabry_solver* h, double* out) {
    DABRY_GUARD(h);
    if (!out) return DABRY_ERR_INVALID;
    E.finish_grid();  // a streamed upload reduces max |w| while it packs
    *out = E.flow_max_norm;
    return DABRY_OK;
}

int dabry_boundary_doubles(dabry_solver* h, int64_t* n) {
    DABRY_GUARD(h);
    if (!n) return DABRY_ERR_INVALID;
    *n = E.boundary_doubles();
    return DABRY_OK;
}

int dabry_boundary_count(dabry_solver* h, int64_t* n) {
    DABRY_GUARD(h);
    if (!n) return DABRY_ERR_INVALID;
    *n = E.boundary_count();
    return DABRY_OK;
}

int dabry_export_boundary(dabry_solver* h, double* packed) {
    DABRY_GUARD(h);
    return E.export_boundary(packed);
}

int dabry_import_boundary(dabry_solver* h, const double* packed) {
    DABRY_GUARD(h);
    return E.import_boundary(packed);
}

int dabry_export_boundary_device(dabry_solver* h, double* d_packed) {
    DABRY_GUARD(h);
    if (!d_packed) return DABRY_ERR_INVALID;
    return E.export_boundary(d_packed, true);
}

int dabry_import_boundary_device(dabry_solver* h, const double* d_packed) {
    DABRY_GUARD(h);
    if (!d_packed) return DABRY_ERR_INVALID;
    return E.import_boundary(d_packed, true);
}

int dabry_get_cost_map_device(dabry_solver* h, int32_t full, double* d_out) {
    DABRY_GUARD(h);
    if (!d_out) return DABRY_ERR_INVALID;
    return E.get_cost_map(full, d_out, true);
}

int dabry_set_cost_map_device(dabry_solver* h, const double* d_map) {
    DABRY_GUARD(h);
    if (!d_map) return DABRY_ERR_INVALID;
    return E.set_cost_map(d_map, true);
}

// ---- in-library collectives ----
int dabry_comm_unique_id(void* id) {
    if (!id) return DABRY_ERR_INVALID;
    return DABRY_BACKEND::comm_unique_id(id, g_comm_error) ? DABRY_OK : DABRY_ERR_UNSUPPORTED;
}

int dabry_comm_create(int32_t device, int32_t rank, int32_t world, const void* id, dabry_comm** out) {
    if (!out || !id || world < 1 || rank < 0 || rank >= world) return DABRY_ERR_INVALID;
    *out = nullptr;
    DABRY_BACKEND::Comm* c = DABRY_BACKEND::comm_create(device, rank, world, id, g_comm_error);
    if (!c) return DABRY_ERR_UNSUPPORTED;
    *out = new dabry_comm{c};
    return DABRY_OK;
}

void dabry_comm_destroy(dabry_comm* c) {
    if (!c) return;
    DABRY_BACKEND::comm_destroy(c->comm);
    delete c;
}

const char* dabry_comm_last_error(void) { return g_comm_error.c_str(); }

int dabry_attach_comm(dabry_solver* h, dabry_comm* c) {
    DABRY_GUARD(h);
    E.comm = c ? c->comm : nullptr;
    return DABRY_OK;
}

int dabry_solve_sharded(dabry_solver* h, int32_t variant, const double* costates) {
    DABRY_GUARD(h);
    return E.solve_sharded(variant, costates);
}

int dabry_get_global_solution(dabry_solver* h, double* cost, int32_t* owner_rank, uint64_t* total_rk_steps) {
    DABRY_GUARD(h);
    if (cost) *cost = E.global_cost;
    if (owner_rank) *owner_rank = E.global_owner;
    if (total_rk_steps) *total_rk_steps = E.global_rk_steps;
    return DABRY_OK;
}

int dabry_time_upper_bound(dabry_solver* h, const double x_init[2], double t_start, double t_end, double out[2]) {
    DABRY_GUARD(h);
    if (!x_init || !out) return DABRY_ERR_INVALID;
    return E.time_upper_bound(x_init, t_start, t_end, out);
}

int dabry_eval_rhs(dabry_solver* h, int64_t n, const double* t, const double* y, double* dy) {
    DABRY_GUARD(h);
    if (n <= 0) return DABRY_OK;
    return E.eval_rhs(n, t, y, dy);
}

int dabry_eval_flow(dabry_solver* h, int64_t n, const double* t, const double* x, double* w, double* jac) {
    DABRY_GUARD(h);
    if (n <= 0) return DABRY_OK;
    return E.eval_flow(n, t, x, w, jac);
}

}  // extern "C"
