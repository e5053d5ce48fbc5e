// capi.cu -- the C ABI (include/stroemung_b200.h) and the host-side tick driver.
//
// Mirrors Simulation::try_from / run_simulation_tick / solve_sor
// (/root/reference/src/simulation.rs:71-99, 239-285, 324-333): the host only sequences
// kernel launches; all field arithmetic happens in the CUDA kernels of this directory.
// The SOR exit test runs on the device (sor_finalize_kernel); passes are enqueued
// speculatively in batches and turn into no-ops once the solve has finished, so the host
// synchronises once per batch instead of once per sweep.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <new>

#include "sor_rb.cuh"

namespace sb {

static thread_local std::string g_error;
static thread_local uint64_t g_err_xy[2] = {0, 0};
static thread_local uint8_t g_err_kind = 0;

void set_error(const std::string &msg) { g_error = msg; }

void prof_mark(sb_sim *s) {
    if (!s->profiling) return;
    if (s->prof_used == s->prof_events.size()) {
        if (s->prof_events.size() >= 8192) return;  // bounded
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        s->prof_events.push_back(e);
    }
    cudaEventRecord(s->prof_events[s->prof_used++], s->stream);
}

// the library's only look at the environment: test hooks and trace switches (DebugKnobs)
static DebugKnobs read_debug_knobs() {
    DebugKnobs k;
    auto geti = [](const char *name, int dflt) {
        const char *e = getenv(name);
        return e ? atoi(e) : dflt;
    };
    k.sor_small = geti("SB_SOR_SMALL", 1) != 0;
    k.sor_mid = geti("SB_SOR_MID", 1) != 0;
    k.sor_mid_ctas = geti("SB_SOR_MID_CTAS", 0);
    k.sor_mid_variant = geti("SB_SOR_MID_VARIANT", 0);
    k.rb_stream = geti("SB_RB_STREAM", 1) != 0;
    k.rb_stream_kinds = geti("SB_RB_STREAM_KINDS", 3);
    k.rb_frozen = geti("SB_RB_FROZEN", 1) != 0;
    if (const char *e = getenv("SB_WALL_WEIGHT")) k.wall_weight = atof(e);
    if (const char *e = getenv("SB_PLAN_OVERHEAD")) k.plan_overhead = atof(e);
    k.trace_plan = getenv("SB_DEBUG_PLAN") != nullptr;
    k.trace_mid = getenv("SB_MID_TRACE") != nullptr;
    k.trace_fin = getenv("SB_FIN_TRACE") != nullptr;
    return k;
}

static bool is_rb(const sb_sim *s) { return s->prm.sor_mode == SB_SOR_RED_BLACK; }

static sb_status validate(const sb_params *p) {
    if (!p || p->nx == 0 || p->ny == 0) {
        set_error("params: nx and ny must be positive");
        return SB_INVALID_ARGUMENT;
    }
    if (p->sor_mode != SB_SOR_REFERENCE_ORDER && p->sor_mode != SB_SOR_RED_BLACK) {
        set_error("params: unknown sor_mode");
        return SB_INVALID_ARGUMENT;
    }
    if (p->temporal_block < 0 || p->temporal_block > 4) {
        set_error("params: temporal_block must be 0..4");
        return SB_INVALID_ARGUMENT;
    }
    if (p->world > 1 && p->sor_mode == SB_SOR_REFERENCE_ORDER) {
        set_error("params: reference-order SOR is single-GPU (the sweep is sequential along x)");
        return SB_INVALID_ARGUMENT;
    }
    return SB_OK;
}

static void destroy(sb_sim *s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    slab_release(s);
    rb_plan_release(s);
    cudaFree(s->d_hist);
    cudaFree(s->d_mid);
    cudaFree(s->d_img);
    cudaFree(s->p[0]); cudaFree(s->p[1]); cudaFree(s->u); cudaFree(s->v); cudaFree(s->f);
    cudaFree(s->gq); cudaFree(s->rhs); cudaFree(s->cflag);
    cudaFree(s->bl.lin); cudaFree(s->bl.ke); cudaFree(s->bl.bu); cudaFree(s->bl.bv);
    cudaFree(s->bl.ru); cudaFree(s->bl.rv); cudaFree(s->bl.nu); cudaFree(s->bl.nv);
    cudaFree(s->bl.wu); cudaFree(s->bl.wv);
    cudaFree(s->d_partial); cudaFree(s->d_scalars); cudaFree(s->d_ctl); cudaFree(s->d_lex_sync); cudaFree(s->d_lex_ll);
    cudaFree(s->d_scan); cudaFree(s->d_err);
    if (s->h_scalars) cudaFreeHost(s->h_scalars);
    if (s->h_ctl) cudaFreeHost(s->h_ctl);
    for (cudaEvent_t e : s->prof_events) cudaEventDestroy(e);
    if (s->ev_sor0) cudaEventDestroy(s->ev_sor0);
    if (s->ev_sor1) cudaEventDestroy(s->ev_sor1);
    if (s->ev_t0) cudaEventDestroy(s->ev_t0);
    if (s->ev_t1) cudaEventDestroy(s->ev_t1);
    for (cudaEvent_t e : s->ev_stage)
        if (e) cudaEventDestroy(e);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

// allocate the handle and its device buffers (zeroed)
static sb_status allocate(const sb_params *params, sb_sim **out) {
    sb_status st = validate(params);
    if (st) return st;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_error("no CUDA device: stroemung_b200 has no CPU fallback");
        return SB_CUDA_ERROR;
    }
    sb_sim *s = new (std::nothrow) sb_sim();
    if (!s) return SB_INVALID_ARGUMENT;
    s->prm = *params;
    s->dbg = read_debug_knobs();
    if (s->prm.temporal_block == 0) s->prm.temporal_block = 4;
    if (params->device >= 0) s->device = params->device;
    else cudaGetDevice(&s->device);
    cudaError_t e = cudaSetDevice(s->device);
    if (e != cudaSuccess) {
        set_error(std::string("cudaSetDevice: ") + cudaGetErrorString(e));
        delete s;
        return SB_CUDA_ERROR;
    }
    Geom &g = s->g;
    g.NX = (int64_t)params->nx;
    g.NY = (int64_t)params->ny;
    int64_t xb = 0, xe = g.NX;
    if (params->world > 1) {
        xb = (int64_t)params->x_begin;
        xe = (int64_t)params->x_end;
        if (xe <= xb || xe > g.NX) {
            set_error("params: bad slab range");
            delete s;
            return SB_INVALID_ARGUMENT;
        }
        if (params->world > SB_MAX_WORLD || params->rank < 0 || params->rank >= params->world) {
            set_error("params: bad rank / world (at most 16 slabs)");
            delete s;
            return SB_INVALID_ARGUMENT;
        }
        if (xe - xb < SB_SLAB_HALO) {
            set_error("params: a slab needs at least 10 owned rows (the halo depth)");
            delete s;
            return SB_INVALID_ARGUMENT;
        }
        s->halo = SB_SLAB_HALO;
    }
    g.nxl = (xe - xb) + 2 * s->halo;
    g.gx0 = xb - s->halo;
    g.own0 = s->halo;
    g.own1 = s->halo + (xe - xb);
    g.pitch = round_up(g.NY, 16);
    s->field_bytes = (size_t)g.nxl * g.pitch * sizeof(double);
    s->flag_bytes = (size_t)g.nxl * g.pitch;
#define SB_TRY(call)                                                            \
    do {                                                                        \
        cudaError_t _e = (call);                                                \
        if (_e != cudaSuccess) {                                                \
            set_error(std::string(#call) + ": " + cudaGetErrorString(_e));     \
            destroy(s);                                                         \
            return SB_CUDA_ERROR;                                               \
        }                                                                       \
    } while (0)
    SB_TRY(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    double **fields[] = {&s->p[0], &s->p[1], &s->u, &s->v, &s->f, &s->gq, &s->rhs};
    for (double **fp : fields) {
        SB_TRY(cudaMalloc(fp, s->field_bytes));
        SB_TRY(cudaMemsetAsync(*fp, 0, s->field_bytes, s->stream));
    }
    SB_TRY(cudaMalloc(&s->cflag, s->flag_bytes));
    SB_TRY(cudaMemsetAsync(s->cflag, 0, s->flag_bytes, s->stream));
    SB_TRY(cudaMalloc(&s->d_scalars, 64 * sizeof(double)));
    SB_TRY(cudaMallocHost(&s->h_scalars, 64 * sizeof(double)));
    SB_TRY(cudaMalloc(&s->d_ctl, 512));
    SB_TRY(cudaMemsetAsync(s->d_ctl, 0, 512, s->stream));
    SB_TRY(cudaMallocHost(&s->h_ctl, 2 * sizeof(SorCtl)));  // [0] solve / read-back, [1] idle image
    memset(s->h_ctl, 0, 2 * sizeof(SorCtl));
    SB_TRY(cudaMalloc(&s->d_err, 2 * sizeof(unsigned long long)));
    SB_TRY(cudaMemcpyAsync(reinterpret_cast<char *>(s->d_ctl) + 256, s->p, 2 * sizeof(double *),
                           cudaMemcpyHostToDevice, s->stream));
    SB_TRY(cudaEventCreate(&s->ev_sor0));
    SB_TRY(cudaEventCreate(&s->ev_sor1));
    SB_TRY(cudaEventCreate(&s->ev_t0));
    SB_TRY(cudaEventCreate(&s->ev_t1));
    for (cudaEvent_t &e : s->ev_stage) SB_TRY(cudaEventCreate(&e));
    if (params->world > 1 && slab_prepare(s) != SB_OK) {
        destroy(s);
        return SB_CUDA_ERROR;
    }
    SB_TRY(cudaStreamSynchronize(s->stream));
#undef SB_TRY
    s->time = params->time;
    s->iterations = params->iterations;
    s->has_initial_norm = params->has_initial_norm;
    s->initial_norm_squared = params->initial_norm_squared;
    *out = s;
    return SB_OK;
}

static sb_status copy_rows_h2d(sb_sim *s, void *dst, const void *src, size_t esize) {
    const Geom &g = s->g;
    int64_t rows = g.own1 - g.own0;
    SB_CUDA(cudaMemcpy2DAsync(static_cast<char *>(dst) + (size_t)g.own0 * g.pitch * esize,
                              g.pitch * esize, src, g.NY * esize, g.NY * esize, rows,
                              cudaMemcpyHostToDevice, s->stream));
    return SB_OK;
}

static sb_status copy_rows_d2h(sb_sim *s, void *dst, const void *src, size_t esize) {
    const Geom &g = s->g;
    int64_t rows = g.own1 - g.own0;
    SB_CUDA(cudaMemcpy2DAsync(dst, g.NY * esize,
                              static_cast<const char *>(src) + (size_t)g.own0 * g.pitch * esize,
                              g.pitch * esize, g.NY * esize, rows, cudaMemcpyDeviceToHost,
                              s->stream));
    return SB_OK;
}

// keep the device control block consistent with the host's view while idle
static sb_status sync_ctl_idle(sb_sim *s) {
    SorCtl *idle = s->h_ctl + 1;  // separate pinned image: the copy below is asynchronous
    SB_CUDA(cudaStreamSynchronize(s->stream));
    memset(idle, 0, sizeof(SorCtl));
    idle->src = s->cur;
    idle->block_T = is_rb(s) ? s->prm.temporal_block : 0;
    SB_CUDA(cudaMemcpyAsync(s->d_ctl, idle, sizeof(SorCtl), cudaMemcpyHostToDevice, s->stream));
    return SB_OK;
}

static sb_status norm_now(sb_sim *s, double *out) {
    sb_status st;
    int nparts = 0;
    if (s->g.NX < 3 || s->g.NY < 3) {
        *out = 0.0 / s->fluid_cells;
        return SB_OK;
    }
    if (is_rb(s)) {
        if ((st = launch_sor_rb_pass(s, &nparts, 1))) return st;
    } else {
        if ((st = launch_norm_partials(s, 0, &nparts))) return st;
    }
    return reduce_norm(s, nparts, out);
}

// Simulation::try_from after the arrays are in place (src/simulation.rs:92-96 and
// src/grid/mod.rs:148-150): classify, ranges, F/G, RHS, initial norm
sb_status finish_create(sb_sim *s) {
    sb_status st;
    if ((st = classify(s))) return st;
    if ((st = sync_ctl_idle(s))) return st;
    if ((st = launch_pressure_range(s))) return st;
    if ((st = launch_speed_range(s))) return st;
    if ((st = launch_fg_rhs(s, 3))) return st;
    if (!s->has_initial_norm) {
        double n = 0.0;
        if ((st = norm_now(s, &n))) return st;
        s->has_initial_norm = 1;
        s->initial_norm_squared = n;
    }
    SB_CUDA(cudaStreamSynchronize(s->stream));
    return SB_OK;
}

// extension A9 (NaSt2D COMP_delt), maxima over Fluid cells; identical in the oracle
static void adapt_delt(sb_sim *s) {
    double dx = s->prm.delx, dy = s->prm.dely;
    double d = (s->prm.reynolds / 2.0) / ((1.0 / (dx * dx)) + (1.0 / (dy * dy)));
    if (s->umax > 0.0) d = fmin(d, dx / s->umax);
    if (s->vmax > 0.0) d = fmin(d, dy / s->vmax);
    s->prm.delt = s->prm.tau * d;
}

// solve_sor (src/simulation.rs:239-285).  test_exit = 0: exactly max_it sweeps.
// cap_hit_out != nullptr: the caller takes over grid.calculate_pressure_range() after a capped
// solve (the tick fuses it into the velocity update) and gets told whether it is due
static sb_status solve(sb_sim *s, uint32_t max_it, int test_exit, uint32_t *iters, double *norm,
                       double *norm_hist_host, int *cap_hit_out = nullptr) {
    sb_status st;
    const bool rb = is_rb(s);
    const int T = rb ? s->prm.temporal_block : 1;
    double *d_hist = nullptr;
    if (norm_hist_host && max_it) {
        if (s->hist_cap < max_it) {
            // grows only; stream-ordered so that no device-wide synchronisation happens while
            // another slab of this process may be waiting for this one
            if (s->d_hist) SB_CUDA(cudaFreeAsync(s->d_hist, s->stream));
            s->d_hist = nullptr;
            s->hist_cap = 0;
            SB_CUDA(cudaMallocAsync(&s->d_hist, (size_t)max_it * sizeof(double), s->stream));
            s->hist_cap = max_it;
        }
        d_hist = s->d_hist;
    }
    if (test_exit && !s->has_initial_norm && max_it > 0 && s->g.NX >= 3 && s->g.NY >= 3) {
        // initial_norm_squared == None (only reachable through sb_set_params: try_from always
        // latches it): get_initial_norm_squared (src/simulation.rs:229-237) then computes and
        // caches the norm AFTER the first sweep (:276), so iteration 0 can only leave through
        // the eps test.  One sweep without the exit rule, latch, test, then the rest.
        uint32_t it1 = 0;
        double n1 = 0.0;
        if ((st = solve(s, 1, 0, &it1, &n1, nullptr))) return st;
        s->has_initial_norm = 1;
        s->initial_norm_squared = n1;
        const double e2 = s->prm.sor_absolute_epsilon * s->prm.sor_absolute_epsilon;
        if (n1 < e2 || max_it == 1) {
            *iters = 1;
            *norm = n1;
            const int capped = !(n1 < e2);
            if (cap_hit_out) *cap_hit_out = capped;
            else if (capped && (st = launch_pressure_range(s))) return st;
            return SB_OK;
        }
        st = solve(s, max_it - 1, 1, iters, norm, nullptr, cap_hit_out);
        *iters += 1;
        return st;
    }
    s->solve_seq++;
    SB_CUDA(cudaEventRecord(s->ev_sor0, s->stream));
    if (max_it == 0 || s->g.NX < 3 || s->g.NY < 3) {
        // no interior: the sweep and the norm are empty loops (0.0 / fluid_cells)
        double n = 0.0;
        uint32_t it = 0;
        if (max_it > 0) {
            // every iteration yields norm 0/fluid_cells; the exit test then decides
            n = 0.0 / s->fluid_cells;
            it = max_it;
            if (test_exit) {
                double eps2 = s->prm.sor_absolute_epsilon * s->prm.sor_absolute_epsilon;
                if ((n < s->initial_norm_squared) || (n < eps2)) it = 1;
            }
            if ((st = launch_pressure_bc(s, 0))) return st;
            for (uint32_t k = 0; norm_hist_host && k < max_it; k++) norm_hist_host[k] = n;
        }
        SB_CUDA(cudaEventRecord(s->ev_sor1, s->stream));
        if (it == max_it && test_exit)
            if ((st = launch_pressure_range(s))) return st;
        *iters = it;
        *norm = n;
        return SB_OK;
    }
    SorCtl *h = s->h_ctl;
    memset(h, 0, sizeof(SorCtl));
    h->active_T = (int32_t)std::min<uint32_t>((uint32_t)T, max_it);
    h->src = s->cur;
    h->max_iterations = max_it;
    h->block_T = rb ? T : 0;
    SB_CUDA(cudaMemcpyAsync(s->d_ctl, h, sizeof(SorCtl), cudaMemcpyHostToDevice, s->stream));
    const double eps2 = s->prm.sor_absolute_epsilon * s->prm.sor_absolute_epsilon;
    const double init = s->initial_norm_squared;
    uint32_t total_passes = (max_it + T - 1) / T + 1;  // +1: a possible shortened redo pass
    uint32_t hint = s->sor_batch_hint ? s->sor_batch_hint : 4;
    uint32_t enq = 0;
    const bool small = rb && max_it > 0 && sor_small_fits(s);
    bool mid = rb && max_it > 0 && !small && !s->mid_unavailable && sor_mid_fits(s);
    if (mid) {
        // a cooperative launch needs all its CTAs resident at once; where the device cannot
        // promise that (SMs taken by another context) the pass kernels take the solve
        st = launch_sor_mid(s, init, eps2, test_exit, d_hist);
        if (st == SB_OK && s->mid_unavailable) mid = false;
        else if (st) return st;
    }
    if (small || mid) {  // the whole solve in one launch (sor_small.cu / sor_mid.cu)
        if (small) {
            s->last_sor_path = 1;
            s->last_sor_ctas = 1;
            if ((st = launch_sor_small(s, init, eps2, test_exit, d_hist))) return st;
        }
        SB_CUDA(cudaMemcpyAsync(h, s->d_ctl, sizeof(SorCtl), cudaMemcpyDeviceToHost, s->stream));
        SB_CUDA(cudaStreamSynchronize(s->stream));
        if (h->pad) {
            set_error("SOR grid barrier timed out (sor_mid.cu)");
            return SB_CUDA_ERROR;
        }
        if (mid && s->last_sor_path == 3 && s->dbg.trace_mid) {
            const double n = std::max(1u, h->iters_done);
            fprintf(stderr, "sor_mid_reg: %u sweeps; cycles per sweep on CTA 0: BC %.0f, red %.0f, "
                            "black %.0f, red residual + publish %.0f, grid barrier %.0f, halo + "
                            "norm %.0f\n", h->iters_done, h->norms[1] / n, h->norms[2] / n,
                    h->norms[3] / n, h->norms[4] / n, h->norms[5] / n, h->norms[6] / n);
        }
    }
    if (!small && !mid) s->last_sor_path = 0;
    while (!small && !mid) {
        uint32_t batch = std::min<uint32_t>(std::max<uint32_t>(hint, 1), 256);
        for (uint32_t b = 0; b < batch; b++) {
            int nparts = 0, fused = 0;
            if (rb) {
                RbFin fin{};
                fin.enabled = 1;
                fin.test_exit = test_exit;
                fin.fluid_cells = s->fluid_cells;
                fin.initial_norm = init;
                fin.eps2 = eps2;
                fin.norm_hist = d_hist;
                if ((st = launch_sor_rb_pass(s, &nparts, 0, &fin, &fused))) return st;
            } else {
                if ((st = launch_pressure_bc(s, 1))) return st;
                if ((st = launch_sor_lex_sweep(s, 1))) return st;
                if ((st = launch_norm_partials(s, 1, &nparts))) return st;
            }
            if (!fused)
                if ((st = launch_sor_finalize(s, nparts, init, eps2, test_exit, d_hist))) return st;
            enq++;
        }
        SB_CUDA(cudaMemcpyAsync(h, s->d_ctl, sizeof(SorCtl), cudaMemcpyDeviceToHost, s->stream));
        SB_CUDA(cudaStreamSynchronize(s->stream));
        if (h->finished) break;
        if (enq > 4 * total_passes + 16) {
            set_error("SOR control block did not finish (internal error)");
            return SB_CUDA_ERROR;
        }
        hint = std::min<uint32_t>(hint * 2, 64);
    }
    SB_CUDA(cudaEventRecord(s->ev_sor1, s->stream));
    s->cur = h->src;
    s->sor_batch_hint = (h->iters_done + T - 1) / T + 1;
    *iters = h->iters_done;
    *norm = h->last_norm;
    if (d_hist) {
        SB_CUDA(cudaMemcpyAsync(norm_hist_host, d_hist, (size_t)h->iters_done * sizeof(double),
                                cudaMemcpyDeviceToHost, s->stream));
        SB_CUDA(cudaStreamSynchronize(s->stream));
    }
    int cap_hit = h->cap_hit;
    if ((st = sync_ctl_idle(s))) return st;
    // grid.calculate_pressure_range() only when the cap is hit (src/simulation.rs:283)
    if (cap_hit_out) *cap_hit_out = cap_hit && test_exit;
    else if (cap_hit && test_exit)
        if ((st = launch_pressure_range(s))) return st;
    return SB_OK;
}

static sb_status tick(sb_sim *s, uint32_t *iters, double *norm) {
    sb_status st;
    if (s->prm.tau > 0.0) {
        // the maxima cached by the last velocity update are stale after an upload of u / v or
        // a re-classification: take them from the fields as they are now
        if (!s->uvmax_valid && (st = launch_speed_range(s))) return st;
        adapt_delt(s);
    }
    cudaEventRecord(s->ev_stage[0], s->stream);
    if ((st = launch_velocity_bc(s))) return st;
    cudaEventRecord(s->ev_stage[1], s->stream);
    if ((st = launch_fg_rhs(s, 3))) return st;
    cudaEventRecord(s->ev_stage[2], s->stream);
    int prange_due = 0;
    if ((st = solve(s, s->prm.max_iterations, 1, iters, norm, nullptr, &prange_due))) return st;
    cudaEventRecord(s->ev_stage[3], s->stream);
    if ((st = launch_adapt_uv(s, prange_due))) return st;
    cudaEventRecord(s->ev_stage[4], s->stream);
    s->time += s->prm.delt;
    s->iterations += 1;
    s->last_sor_iterations = *iters;
    s->last_norm_squared = *norm;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, s->ev_sor0, s->ev_sor1) == cudaSuccess) s->last_sor_ms = ms;
    return slab_check_error(s);
}

}  // namespace sb

using namespace sb;

#define SB_ENTER(s)                                   \
    if (!(s)) {                                       \
        set_error("null handle");                     \
        return SB_INVALID_ARGUMENT;                   \
    }                                                 \
    if ((s)->slab && !(s)->connected) {               \
        set_error("slab handle: call sb_slab_connect first"); \
        return SB_INVALID_ARGUMENT;                   \
    }                                                 \
    SB_CUDA(cudaSetDevice((s)->device))

extern "C" {

sb_status sb_create(const sb_params *params, const double *p, const double *u, const double *v,
                    const uint8_t *kind, const sb_boundary_velocity *velocities,
                    size_t n_velocities, sb_sim **out) {
    if (!out) return SB_INVALID_ARGUMENT;
    *out = nullptr;
    if (!kind) {
        set_error("sb_create: kind must not be NULL");
        return SB_INVALID_ARGUMENT;
    }
    sb_sim *s = nullptr;
    sb_status st = allocate(params, &s);
    if (st) return st;
    auto fail = [&](sb_status code) {
        if (code == SB_BOUNDARY_TOO_THIN) {
            g_err_xy[0] = s->err_xy[0];
            g_err_xy[1] = s->err_xy[1];
            g_err_kind = s->err_kind;
        }
        destroy(s);
        return code;
    };
    if (p && (st = copy_rows_h2d(s, s->p[0], p, 8))) return fail(st);
    if (u && (st = copy_rows_h2d(s, s->u, u, 8))) return fail(st);
    if (v && (st = copy_rows_h2d(s, s->v, v, 8))) return fail(st);
    // kinds go through p[1] as a staging area (same pitch in bytes is not needed: use cflag)
    if ((st = copy_rows_h2d(s, s->cflag, kind, 1))) return fail(st);
    if ((st = launch_mark_valid(s, nullptr, 0))) return fail(st);
    if (velocities && n_velocities) s->velocities.assign(velocities, velocities + n_velocities);
    // a slab finishes construction collectively in sb_slab_connect
    if (!s->slab && (st = finish_create(s))) return fail(st);
    if (cudaStreamSynchronize(s->stream) != cudaSuccess) {
        set_error("sb_create: the construction kernels failed");
        return fail(SB_CUDA_ERROR);
    }
    *out = s;
    return SB_OK;
}

sb_status sb_create_preset(const sb_params *params, int32_t preset, const double *args,
                           size_t n_args, sb_sim **out) {
    if (!out) return SB_INVALID_ARGUMENT;
    *out = nullptr;
    if (preset < 0 || preset > 5) {
        set_error("sb_create_preset: unknown preset");
        return SB_INVALID_ARGUMENT;
    }
    if (n_args && !args) {
        set_error("sb_create_preset: args is NULL");
        return SB_INVALID_ARGUMENT;
    }
    sb_sim *s = nullptr;
    sb_status st = allocate(params, &s);
    if (st) return st;
    double a[3] = {0, 0, 0};
    for (size_t i = 0; i < n_args && i < 3; i++) a[i] = args[i];
    const int64_t NX = s->g.NX, NY = s->g.NY;
    double lid_u = 1.0;
    if (preset == 5) {
        lid_u = n_args >= 1 ? args[0] : 1.0;
        a[0] = a[1] = a[2] = 0;
    }
    auto fail = [&](sb_status code) {
        if (code == SB_BOUNDARY_TOO_THIN) {
            g_err_xy[0] = s->err_xy[0];
            g_err_xy[1] = s->err_xy[1];
            g_err_kind = s->err_kind;
        }
        destroy(s);
        return code;
    };
    if ((st = launch_preset(s, preset, a))) return fail(st);
    // boundary velocities of the preset (Inflow [1, 0], src/grid/presets.rs:26-28)
    if (preset >= 1 && preset <= 4) {
        int64_t y_end = preset == 4 ? std::min<int64_t>((int64_t)a[1], NY - 1) : NY - 1;
        for (int64_t y = 1; y < y_end; y++)
            s->velocities.push_back(sb_boundary_velocity{0, (uint64_t)y, 1.0, 0.0});
    } else if (preset == 5) {
        for (int64_t x = 1; x <= NX - 2; x++)
            s->velocities.push_back(sb_boundary_velocity{(uint64_t)x, 0, lid_u, 0.0});
    }
    if (!s->slab && (st = finish_create(s))) return fail(st);
    if (cudaStreamSynchronize(s->stream) != cudaSuccess) {
        set_error("sb_create: the construction kernels failed");
        return fail(SB_CUDA_ERROR);
    }
    *out = s;
    return SB_OK;
}

void sb_destroy(sb_sim *sim) {
    if (sim && sim->slab && sim->dbg.trace_fin) {
        cudaSetDevice(sim->device);
        cudaStreamSynchronize(sim->stream);
        sb::dump_finalize_trace(sim->link.rank);
    }
    destroy(sim);
}

sb_status sb_tick(sb_sim *sim, uint32_t *sor_iterations, double *norm_squared) {
    SB_ENTER(sim);
    uint32_t it = 0;
    double n = 0.0;
    sb_status st = tick(sim, &it, &n);
    if (st) return st;
    if (sor_iterations) *sor_iterations = it;
    if (norm_squared) *norm_squared = n;
    return SB_OK;
}

static double *field_ptr(sb_sim *s, sb_field f);

sb_status sb_tick_host(sb_sim *sim, const double *p_in, const double *u_in, const double *v_in,
                       double *p_out, double *u_out, double *v_out, uint32_t *sor_iterations,
                       double *norm_squared) {
    SB_ENTER(sim);
    if (!p_in || !u_in || !v_in || !p_out || !u_out || !v_out) return SB_INVALID_ARGUMENT;
    if (sim->slab) {
        set_error("sb_tick_host: row slabs upload with sb_upload + sb_slab_sync_halos");
        return SB_INVALID_ARGUMENT;
    }
    sb_status st;
    const double *in[3] = {p_in, u_in, v_in};
    double *out[3] = {p_out, u_out, v_out};
    const sb_field fld[3] = {SB_FIELD_P, SB_FIELD_U, SB_FIELD_V};
    sim->uvmax_valid = false;
    for (int i = 0; i < 3; i++)
        if ((st = copy_rows_h2d(sim, field_ptr(sim, fld[i]), in[i], 8))) return st;
    uint32_t it = 0;
    double n = 0.0;
    if ((st = tick(sim, &it, &n))) return st;
    for (int i = 0; i < 3; i++)   // field_ptr again: the tick may have swapped the p buffers
        if ((st = copy_rows_d2h(sim, out[i], field_ptr(sim, fld[i]), 8))) return st;
    SB_CUDA(cudaStreamSynchronize(sim->stream));
    if (sor_iterations) *sor_iterations = it;
    if (norm_squared) *norm_squared = n;
    return SB_OK;
}

sb_status sb_run_ticks(sb_sim *sim, uint32_t n, uint32_t *sor_iterations, double *norm_squared) {
    SB_ENTER(sim);
    uint32_t it = 0;
    double nn = 0.0;
    for (uint32_t k = 0; k < n; k++) {
        sb_status st = tick(sim, &it, &nn);
        if (st) return st;
    }
    if (sor_iterations) *sor_iterations = it;
    if (norm_squared) *norm_squared = nn;
    return SB_OK;
}

sb_status sb_set_boundary_u_and_v(sb_sim *sim) {
    SB_ENTER(sim);
    sb_status st = launch_velocity_bc(sim);
    if (st) return st;
    SB_CUDA(cudaStreamSynchronize(sim->stream));
    return SB_OK;
}

sb_status sb_calculate_f_and_g(sb_sim *sim) {
    SB_ENTER(sim);
    sb_status st = launch_fg(sim);
    if (st) return st;
    SB_CUDA(cudaStreamSynchronize(sim->stream));
    return SB_OK;
}

sb_status sb_calculate_rhs(sb_sim *sim) {
    SB_ENTER(sim);
    sb_status st = launch_rhs(sim);
    if (st) return st;
    SB_CUDA(cudaStreamSynchronize(sim->stream));
    return SB_OK;
}

sb_status sb_copy_pressure_to_boundaries(sb_sim *sim) {
    SB_ENTER(sim);
    sb_status st = launch_pressure_bc(sim, 0);
    if (st) return st;
    SB_CUDA(cudaStreamSynchronize(sim->stream));
    return SB_OK;
}

sb_status sb_calculate_norm_squared(sb_sim *sim, double *norm_squared) {
    SB_ENTER(sim);
    double n = 0.0;
    sb_status st = norm_now(sim, &n);
    if (st) return st;
    if (norm_squared) *norm_squared = n;
    return SB_OK;
}

sb_status sb_solve_sor(sb_sim *sim, uint32_t *sor_iterations, double *norm_squared) {
    SB_ENTER(sim);
    uint32_t it = 0;
    double n = 0.0;
    sb_status st = solve(sim, sim->prm.max_iterations, 1, &it, &n, nullptr);
    if (st) return st;
    sim->last_sor_iterations = it;
    sim->last_norm_squared = n;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, sim->ev_sor0, sim->ev_sor1) == cudaSuccess) sim->last_sor_ms = ms;
    if (sor_iterations) *sor_iterations = it;
    if (norm_squared) *norm_squared = n;
    return SB_OK;
}

sb_status sb_sor_sweeps(sb_sim *sim, uint32_t n, double *norms) {
    SB_ENTER(sim);
    uint32_t it = 0;
    double nn = 0.0;
    sb_status st = solve(sim, n, 0, &it, &nn, norms);
    if (st) return st;
    SB_CUDA(cudaStreamSynchronize(sim->stream));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, sim->ev_sor0, sim->ev_sor1) == cudaSuccess) sim->last_sor_ms = ms;
    sim->last_sor_iterations = it;
    sim->last_norm_squared = nn;
    return SB_OK;
}

sb_status sb_set_u_and_v(sb_sim *sim) {
    SB_ENTER(sim);
    return launch_adapt_uv(sim);
}

sb_status sb_calculate_pressure_range(sb_sim *sim) {
    SB_ENTER(sim);
    return launch_pressure_range(sim);
}

sb_status sb_calculate_speed_range(sb_sim *sim) {
    SB_ENTER(sim);
    return launch_speed_range(sim);
}

// render_simulation (src/visualization.rs:79-105): the frame of the owned rows, RGBA8,
// width = owned rows, height = ny, into host memory
sb_status sb_render_rgba(sb_sim *sim, int32_t color_type, uint8_t *dst) {
    SB_ENTER(sim);
    if (!dst || (color_type != SB_COLOR_PRESSURE && color_type != SB_COLOR_SPEED)) {
        set_error("sb_render_rgba: dst is NULL or unknown color_type");
        return SB_INVALID_ARGUMENT;
    }
    const size_t bytes = (size_t)(sim->g.own1 - sim->g.own0) * (size_t)sim->g.NY * 4;
    if (!sim->d_img) SB_CUDA(cudaMallocAsync(&sim->d_img, bytes, sim->stream));
    sb_status st = launch_render(sim, color_type == SB_COLOR_SPEED, sim->d_img);
    if (st) return st;
    SB_CUDA(cudaMemcpyAsync(dst, sim->d_img, bytes, cudaMemcpyDeviceToHost, sim->stream));
    SB_CUDA(cudaStreamSynchronize(sim->stream));
    return SB_OK;
}

static double *field_ptr(sb_sim *s, sb_field f) {
    switch (f) {
    case SB_FIELD_P: return s->p[s->cur];
    case SB_FIELD_U: return s->u;
    case SB_FIELD_V: return s->v;
    case SB_FIELD_F: return s->f;
    case SB_FIELD_G: return s->gq;
    case SB_FIELD_RHS: return s->rhs;
    default: return nullptr;
    }
}

sb_status sb_download(sb_sim *sim, sb_field field, void *dst) {
    SB_ENTER(sim);
    if (!dst) return SB_INVALID_ARGUMENT;
    sb_status st;
    if (field == SB_FIELD_KIND || field == SB_FIELD_EDGE) {
        if ((st = copy_rows_d2h(sim, dst, sim->cflag, 1))) return st;
        SB_CUDA(cudaStreamSynchronize(sim->stream));
        uint8_t *b = static_cast<uint8_t *>(dst);
        size_t n = (size_t)(sim->g.own1 - sim->g.own0) * sim->g.NY;
        if (field == SB_FIELD_KIND) for (size_t i = 0; i < n; i++) b[i] = (uint8_t)cf_kind(b[i]);
        else for (size_t i = 0; i < n; i++) b[i] = (uint8_t)cf_edge(b[i]);
        return SB_OK;
    }
    double *src = field_ptr(sim, field);
    if (!src) {
        set_error("sb_download: unknown field");
        return SB_INVALID_ARGUMENT;
    }
    if ((st = copy_rows_d2h(sim, dst, src, 8))) return st;
    SB_CUDA(cudaStreamSynchronize(sim->stream));
    return SB_OK;
}

sb_status sb_upload(sb_sim *sim, sb_field field, const void *src) {
    SB_ENTER(sim);
    if (!src) return SB_INVALID_ARGUMENT;
    sb_status st;
    if (field == SB_FIELD_KIND) {
        // new kinds, old edge classes: like the reference, the boundary list is stale until
        // sb_rebuild_boundary_list (src/lib.rs:60-70)
        uint8_t *tmp = nullptr;
        SB_CUDA(cudaMallocAsync(&tmp, sim->flag_bytes, sim->stream));
        SB_CUDA(cudaMemsetAsync(tmp, 0, sim->flag_bytes, sim->stream));
        if ((st = copy_rows_h2d(sim, tmp, src, 1))) { cudaFreeAsync(tmp, sim->stream); return st; }
        st = launch_mark_valid(sim, tmp, 1);
        cudaFreeAsync(tmp, sim->stream);
        cudaStreamSynchronize(sim->stream);
        return st;
    }
    if (field == SB_FIELD_EDGE) {
        set_error("sb_upload: the edge class is derived (read-only)");
        return SB_INVALID_ARGUMENT;
    }
    double *dst = field_ptr(sim, field);
    if (!dst) return SB_INVALID_ARGUMENT;
    if (field == SB_FIELD_U || field == SB_FIELD_V) sim->uvmax_valid = false;
    if ((st = copy_rows_h2d(sim, dst, src, 8))) return st;
    SB_CUDA(cudaStreamSynchronize(sim->stream));
    return SB_OK;
}

void *sb_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) return nullptr;
    return p;
}

void sb_host_free(void *ptr) {
    if (ptr) cudaFreeHost(ptr);
}

sb_status sb_get_state(sb_sim *sim, sb_state *st) {
    if (!sim || !st) return SB_INVALID_ARGUMENT;
    memset(st, 0, sizeof(*st));
    st->time = sim->time;
    st->delt = sim->prm.delt;
    st->iterations = sim->iterations;
    st->has_initial_norm = sim->has_initial_norm;
    st->initial_norm_squared = sim->initial_norm_squared;
    st->pressure_range[0] = sim->pressure_range[0];
    st->pressure_range[1] = sim->pressure_range[1];
    st->speed_range[0] = sim->speed_range[0];
    st->speed_range[1] = sim->speed_range[1];
    st->fluid_cells = sim->fluid_cells;
    st->n_boundary = sim->bl.n;
    st->last_sor_iterations = sim->last_sor_iterations;
    st->last_norm_squared = sim->last_norm_squared;
    return SB_OK;
}

sb_status sb_set_params(sb_sim *sim, const sb_params *p) {
    SB_ENTER(sim);
    if (!p) return SB_INVALID_ARGUMENT;
    if (p->sor_mode != SB_SOR_REFERENCE_ORDER && p->sor_mode != SB_SOR_RED_BLACK)
        return SB_INVALID_ARGUMENT;
    if (p->temporal_block < 0 || p->temporal_block > 4) return SB_INVALID_ARGUMENT;
    if (sim->prm.world > 1 && p->sor_mode == SB_SOR_REFERENCE_ORDER) return SB_INVALID_ARGUMENT;
    sim->prm.delx = p->delx; sim->prm.dely = p->dely;
    sim->prm.delt = p->delt; sim->prm.gamma = p->gamma; sim->prm.reynolds = p->reynolds;
    sim->prm.sor_absolute_epsilon = p->sor_absolute_epsilon; sim->prm.omega = p->omega;
    sim->prm.max_iterations = p->max_iterations; sim->prm.tau = p->tau;
    sim->prm.sor_mode = p->sor_mode;
    sim->prm.temporal_block = p->temporal_block ? p->temporal_block : 4;
    sim->time = p->time;
    sim->iterations = p->iterations;
    sim->has_initial_norm = p->has_initial_norm;
    sim->initial_norm_squared = p->initial_norm_squared;
    sim->sor_batch_hint = 0;
    return sync_ctl_idle(sim);
}

sb_status sb_set_boundary_velocities(sb_sim *sim, const sb_boundary_velocity *v, size_t n) {
    SB_ENTER(sim);
    sim->velocities.clear();
    if (v && n) sim->velocities.assign(v, v + n);
    return apply_velocity_table(sim);
}

sb_status sb_get_boundary_velocities(sb_sim *sim, sb_boundary_velocity *v, size_t capacity,
                                     size_t *n) {
    if (!sim || !n) return SB_INVALID_ARGUMENT;
    *n = sim->velocities.size();
    if (v) std::copy_n(sim->velocities.begin(), std::min(capacity, sim->velocities.size()), v);
    return SB_OK;
}

sb_status sb_rebuild_boundary_list(sb_sim *sim) {
    SB_ENTER(sim);
    sb_status st;
    if (sim->slab && (st = slab_sync_halos(sim, 1))) return st;  // kinds of the halo rows
    st = classify(sim);
    if (st == SB_BOUNDARY_TOO_THIN) {
        g_err_xy[0] = sim->err_xy[0];
        g_err_xy[1] = sim->err_xy[1];
        g_err_kind = sim->err_kind;
        // the previous list stays active (src/grid/mod.rs:232-233): put its edge classes back
        sb_status st2 = restore_edges_from_list(sim);
        if (st2) return st2;
        cudaStreamSynchronize(sim->stream);
    }
    return st;
}

sb_status sb_boundary_list(sb_sim *sim, uint64_t *index, uint8_t *edge, uint64_t capacity,
                           uint64_t *n) {
    SB_ENTER(sim);
    if (n) *n = sim->bl.n;
    uint64_t m = std::min<uint64_t>(capacity, sim->bl.n);
    if (m == 0) return SB_OK;
    std::vector<int64_t> lin(m);
    std::vector<uint8_t> ke(m);
    SB_CUDA(cudaMemcpyAsync(lin.data(), sim->bl.lin, m * sizeof(int64_t), cudaMemcpyDeviceToHost,
                            sim->stream));
    SB_CUDA(cudaMemcpyAsync(ke.data(), sim->bl.ke, m, cudaMemcpyDeviceToHost, sim->stream));
    SB_CUDA(cudaStreamSynchronize(sim->stream));
    for (uint64_t k = 0; k < m; k++) {
        int64_t lx = lin[k] / sim->g.pitch, y = lin[k] - lx * sim->g.pitch;
        if (index) index[k] = (uint64_t)((sim->g.gx0 + lx) * sim->g.NY + y);
        if (edge) edge[k] = ke[k] >> 3;
    }
    return SB_OK;
}

sb_status sb_edit_cells(sb_sim *sim, uint64_t x, uint64_t y, uint8_t kind, double bu, double bv,
                        int32_t *applied) {
    SB_ENTER(sim);
    if (kind > SB_KIND_MOVING_WALL) return SB_INVALID_ARGUMENT;
    if (sim->slab) {
        set_error("sb_edit_cells: interactive edits are single-GPU (the GUI path)");
        return SB_INVALID_ARGUMENT;
    }
    sb_status st;
    double *d_backup = sim->d_scalars + 8;  // 20 doubles + 1 flag word
    int32_t *d_mod = reinterpret_cast<int32_t *>(sim->d_scalars + 32);
    if ((st = launch_edit_block(sim, (int64_t)x, (int64_t)y, kind, d_backup, 0, d_mod))) return st;
    int32_t mod = 0;
    SB_CUDA(cudaMemcpyAsync(&mod, d_mod, sizeof(mod), cudaMemcpyDeviceToHost, sim->stream));
    SB_CUDA(cudaStreamSynchronize(sim->stream));
    if (applied) *applied = mod;
    if (!mod) return SB_OK;
    std::vector<sb_boundary_velocity> saved = sim->velocities;
    if (kind == SB_KIND_INFLOW || kind == SB_KIND_MOVING_WALL) {
        // one table entry per cell: drop older entries of the block first
        auto &tab = sim->velocities;
        tab.erase(std::remove_if(tab.begin(), tab.end(),
                                 [&](const sb_boundary_velocity &e) {
                                     return e.x >= x && e.x <= x + 1 && e.y >= y && e.y <= y + 1;
                                 }),
                  tab.end());
        for (int k = 0; k < 4; k++)
            tab.push_back(sb_boundary_velocity{x + (k & 1), y + (k >> 1), bu, bv});
    }
    st = classify(sim);
    if (st == SB_BOUNDARY_TOO_THIN) {
        // roll the four cells back; the old list (still in force) matches them again
        g_err_xy[0] = sim->err_xy[0];
        g_err_xy[1] = sim->err_xy[1];
        g_err_kind = sim->err_kind;
        sim->velocities = saved;
        sb_status st2 = launch_edit_block(sim, (int64_t)x, (int64_t)y, kind, d_backup, 1, d_mod);
        if (st2) return st2;
        if ((st2 = restore_edges_from_list(sim))) return st2;
        SB_CUDA(cudaStreamSynchronize(sim->stream));
        if (applied) *applied = 0;
        return SB_OK;
    }
    return st;
}

sb_status sb_error_cell(const sb_sim *sim, uint64_t xy[2], uint8_t *kind) {
    if (xy) {
        xy[0] = sim ? sim->err_xy[0] : g_err_xy[0];
        xy[1] = sim ? sim->err_xy[1] : g_err_xy[1];
    }
    if (kind) *kind = sim ? sim->err_kind : g_err_kind;
    return SB_OK;
}

const char *sb_last_error_string(void) { return g_error.c_str(); }

static sb_status cellop(int op, const double *u, const double *v, double s0, double s1, double s2,
                        double s3, double s4, double *out) {
    if (!out) return SB_INVALID_ARGUMENT;
    double sc[5] = {s0, s1, s2, s3, s4};
    return launch_cellop(op, u, v, sc, out);
}

sb_status sb_du2dx(const double u[9], double delx, double gamma, double *out) {
    return cellop(0, u, nullptr, delx, gamma, 0, 0, 0, out);
}
sb_status sb_duvdx(const double u[9], const double v[9], double delx, double gamma, double *out) {
    return cellop(1, u, v, delx, gamma, 0, 0, 0, out);
}
sb_status sb_duvdy(const double u[9], const double v[9], double dely, double gamma, double *out) {
    return cellop(2, u, v, dely, gamma, 0, 0, 0, out);
}
sb_status sb_dv2dy(const double v[9], double dely, double gamma, double *out) {
    return cellop(3, nullptr, v, dely, gamma, 0, 0, 0, out);
}
sb_status sb_laplacian(const double e[9], double delx, double dely, double *out) {
    return cellop(4, e, nullptr, delx, dely, 0, 0, 0, out);
}
sb_status sb_residual(const double p[9], double delx, double dely, double rhs, double *out) {
    return cellop(5, p, nullptr, delx, dely, rhs, 0, 0, out);
}
sb_status sb_calculate_f(const double u[9], const double v[9], double delx, double dely,
                         double delt, double gamma, double reynolds, double *out) {
    return cellop(6, u, v, delx, dely, delt, gamma, reynolds, out);
}
sb_status sb_calculate_g(const double u[9], const double v[9], double delx, double dely,
                         double delt, double gamma, double reynolds, double *out) {
    return cellop(7, u, v, delx, dely, delt, gamma, reynolds, out);
}

sb_status sb_profile_enable(sb_sim *sim, int32_t enable) {
    SB_ENTER(sim);
    SB_CUDA(cudaStreamSynchronize(sim->stream));
    sim->profiling = enable != 0;
    sim->prof_used = 0;
    return SB_OK;
}

sb_status sb_profile_read(sb_sim *sim, double *ms, size_t capacity, size_t *n) {
    SB_ENTER(sim);
    SB_CUDA(cudaStreamSynchronize(sim->stream));
    size_t pairs = sim->prof_used / 2;
    if (n) *n = pairs;
    for (size_t i = 0; i < pairs && i < capacity && ms; i++) {
        float t = 0.f;
        SB_CUDA(cudaEventElapsedTime(&t, sim->prof_events[2 * i], sim->prof_events[2 * i + 1]));
        ms[i] = t;
    }
    sim->prof_used = 0;
    return SB_OK;
}

sb_status sb_timer_begin(sb_sim *sim) {
    SB_ENTER(sim);
    SB_CUDA(cudaStreamSynchronize(sim->stream));
    SB_CUDA(cudaEventRecord(sim->ev_t0, sim->stream));
    return SB_OK;
}

sb_status sb_timer_end(sb_sim *sim, double *elapsed_ms) {
    SB_ENTER(sim);
    SB_CUDA(cudaEventRecord(sim->ev_t1, sim->stream));
    SB_CUDA(cudaEventSynchronize(sim->ev_t1));
    float ms = 0.f;
    SB_CUDA(cudaEventElapsedTime(&ms, sim->ev_t0, sim->ev_t1));
    if (elapsed_ms) *elapsed_ms = ms;
    return SB_OK;
}

sb_status sb_rb_plan(const sb_sim *sim, int32_t *tile_kernel_tiles, int32_t *stream_items) {
    if (!sim || !tile_kernel_tiles || !stream_items) return SB_INVALID_ARGUMENT;
    *tile_kernel_tiles = sim->plan.n_slow;
    *stream_items = sim->plan.n_items;
    return SB_OK;
}
uint64_t sb_kernel_launches(const sb_sim *sim) { return sim ? sim->launches : 0; }

int32_t sb_last_sor_path(const sb_sim *sim, int32_t *ctas) {
    if (!sim) return 0;
    if (ctas) *ctas = sim->last_sor_path ? sim->last_sor_ctas : 0;
    return sim->last_sor_path;
}
double sb_last_sor_ms(const sb_sim *sim) { return sim ? sim->last_sor_ms : 0.0; }
sb_status sb_last_stage_ms(sb_sim *sim, double ms[4]) {
    SB_ENTER(sim);
    if (!ms) return SB_INVALID_ARGUMENT;
    SB_CUDA(cudaStreamSynchronize(sim->stream));
    for (int i = 0; i < 4; i++) {
        float t = 0.f;
        ms[i] = cudaEventElapsedTime(&t, sim->ev_stage[i], sim->ev_stage[i + 1]) == cudaSuccess ? t : 0.0;
    }
    cudaGetLastError();   // events never recorded (no tick yet): zeros, not an error
    return SB_OK;
}
void *sb_stream(const sb_sim *sim) { return sim ? (void *)sim->stream : nullptr; }
const char *sb_version(void) { return "stroemung_b200 0.1.0 (sm_100a)"; }

}  // extern "C"
