/*
 * stroemung_b200_pipeline.hpp -- ticks on HOST arrays, several in flight (C++ host mirror).
 *
 * The reference keeps `pressure`, `u`, `v` in host arrays (src/grid/mod.rs:112-125) and
 * `run_simulation_tick` (src/simulation.rs:324-333) updates them in place.  A host-side owner of
 * the fields that ticks them through the GPU pays three fields up and three fields down per
 * tick (8192^2: 1.6 GB each way, ~29 ms per direction on PCIe 5 x16 against a 10 ms tick), and
 * one handle does upload -> tick -> download in series.  PCIe is full duplex and the copy
 * engines run beside the SMs, so `HostPipeline` keeps `depth` handles of the same geometry,
 * each with its own CUDA stream, driven by its own host thread through `sb_tick_host`: the
 * upload of one request overlaps the kernels of a second and the download of a third.  This is
 * the end-to-end path bench.py times (its `e2e` leg drives the same C-ABI call from Python
 * threads, stroemung_b200/pipeline.py).
 *
 *     HostPipeline pipe(3, [&] { return Simulation::try_from(unfinalized, ext); });
 *     PinnedField p(pipe.field_len()), u(...), v(...);        // page-locked staging
 *     auto fut = pipe.submit(p.data(), u.data(), v.data());   // in place
 *     auto [sor_iterations, norm_squared] = fut.get();        // p, u, v hold the next level
 *
 * A request carries the fields only: time, iteration count and the latched
 * `initial_norm_squared` of the exit rule (src/simulation.rs:229-237, :279) are state of the
 * handles -- build them with an explicit `initial_norm_squared` when the exit rule matters.
 * Requests are independent simulations of one geometry (ensembles, sweeps over initial states);
 * a single simulation stays on the device and uses `Simulation::run_ticks`, which moves nothing.
 */
#ifndef STROEMUNG_B200_PIPELINE_HPP
#define STROEMUNG_B200_PIPELINE_HPP

#include <condition_variable>
#include <deque>
#include <functional>
#include <future>
#include <mutex>
#include <thread>
#include <vector>

#include "stroemung_b200.hpp"

namespace stroemung {

/* one field of page-locked host memory (sb_host_alloc): makes the copies truly asynchronous */
class PinnedField {
  public:
    explicit PinnedField(std::size_t len) : len_(len), p_(static_cast<Real *>(sb_host_alloc(len * sizeof(Real)))) {
        if (!p_) throw CudaError("sb_host_alloc failed");
    }
    ~PinnedField() { sb_host_free(p_); }
    PinnedField(const PinnedField &) = delete;
    PinnedField &operator=(const PinnedField &) = delete;
    PinnedField(PinnedField &&o) noexcept : len_(o.len_), p_(o.p_) { o.p_ = nullptr; }
    Real *data() { return p_; }
    const Real *data() const { return p_; }
    std::size_t len() const { return len_; }
    Real &operator[](std::size_t i) { return p_[i]; }
    const Real &operator[](std::size_t i) const { return p_[i]; }

  private:
    std::size_t len_;
    Real *p_;
};

class HostPipeline {
  public:
    using Result = std::pair<std::uint32_t, Real>; /* (sor_iterations, norm_squared) */

    /* `make_sim()` builds one single-GPU Simulation; it is called `depth` times */
    HostPipeline(std::size_t depth, const std::function<Simulation()> &make_sim) {
        if (depth < 1) throw InvalidArgument("HostPipeline: depth must be >= 1");
        sims_.reserve(depth);
        for (std::size_t i = 0; i < depth; ++i) sims_.push_back(make_sim());
        for (std::size_t i = 0; i < depth; ++i) workers_.emplace_back([this, i] { work(sims_[i]); });
    }
    ~HostPipeline() { close(); }
    HostPipeline(const HostPipeline &) = delete;
    HostPipeline &operator=(const HostPipeline &) = delete;

    std::size_t depth() const { return sims_.size(); }
    std::size_t field_len() const { return sims_.empty() ? 0 : sims_[0].size[0] * sims_[0].size[1]; }

    /* one tick of the state (p, u, v), [nx][ny] f64 each; outputs default to the inputs (in
     * place, like the reference's tick on its own arrays).  The future yields the tick's pair
     * once the outputs hold the new state, or rethrows the SimulationError of the tick. */
    std::future<Result> submit(Real *p, Real *u, Real *v, Real *p_out = nullptr, Real *u_out = nullptr,
                               Real *v_out = nullptr) {
        Job job{p, u, v, p_out ? p_out : p, u_out ? u_out : u, v_out ? v_out : v, {}};
        std::future<Result> fut = job.done.get_future();
        {
            std::lock_guard<std::mutex> lock(m_);
            if (closed_) throw InvalidArgument("HostPipeline: submit after close");
            q_.push_back(std::move(job));
        }
        cv_.notify_one();
        return fut;
    }

    /* finish what is queued, stop the threads, destroy the handles */
    void close() {
        {
            std::lock_guard<std::mutex> lock(m_);
            if (closed_) return;
            closed_ = true;
        }
        cv_.notify_all();
        for (std::thread &t : workers_) t.join();
        workers_.clear();
        sims_.clear();
    }

  private:
    struct Job {
        const Real *p, *u, *v;
        Real *p_out, *u_out, *v_out;
        std::promise<Result> done;
    };
    std::vector<Simulation> sims_;
    std::vector<std::thread> workers_;
    std::deque<Job> q_;
    std::mutex m_;
    std::condition_variable cv_;
    bool closed_ = false;

    void work(Simulation &sim) {
        for (;;) {
            Job job;
            {
                std::unique_lock<std::mutex> lock(m_);
                cv_.wait(lock, [this] { return closed_ || !q_.empty(); });
                if (q_.empty()) return; /* closed and drained */
                job = std::move(q_.front());
                q_.pop_front();
            }
            try {
                job.done.set_value(sim.tick_host(job.p, job.u, job.v, job.p_out, job.u_out, job.v_out));
            } catch (...) {
                job.done.set_exception(std::current_exception());
            }
        }
    }
};

} // namespace stroemung

#endif /* STROEMUNG_B200_PIPELINE_HPP */
