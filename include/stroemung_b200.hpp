/*
 * stroemung_b200.hpp -- C++17 host-side mirror of the reference's Rust API over the C ABI.
 *
 * The reference (wickedchicken/stroemung 0.1.2) is compiled Rust and this image has no
 * rustc, so the host layer above include/stroemung_b200.h is written in C++ with the
 * reference's own names, argument meaning and error behaviour:
 *
 *   stroemung::Simulation          src/simulation.rs:49-69   (try_from :71-99, tick :324-333)
 *   stroemung::SimulationGrid      src/grid/mod.rs:112-125   (pub fns :202-268, :343-651)
 *   stroemung::UnfinalizedSimulation[Grid]  src/simulation.rs:30-44, src/grid/mod.rs:85-92
 *   stroemung::Cell / BoundaryCell src/cell.rs:6-23
 *   stroemung::EdgeType            src/grid/mod.rs:19-49
 *   stroemung::BoundaryList        src/grid/mod.rs:61-69
 *   stroemung::presets::*          src/grid/presets.rs:8-87
 *   stroemung::math::* , calculate_f / calculate_g   src/math.rs:19-186, src/simulation.rs:349-392
 *
 * State lives on the GPU.  A Rust `pub` array field becomes a pair of methods: `sim.f()`
 * downloads, `grid.set_pressure(a)` uploads.  `Result<T, E>` becomes `T` or a thrown
 * `SimulationError` whose `what()` is the reference's `thiserror` message.  Nothing in this
 * header computes: every method is one call into libstroemung_b200.so, and there is no CPU
 * fallback -- without a CUDA device `Simulation::try_from` throws `CudaError`.
 *
 * Header-only; link with -lstroemung_b200.  tests/cpp/ replays the reference's own tests
 * through this header (tests/test_cpp_mirror.py builds and runs them).
 */
#ifndef STROEMUNG_B200_HPP
#define STROEMUNG_B200_HPP

#include <array>
#include <cstddef>
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "stroemung_b200.h"

namespace stroemung {

/* src/math.rs:3, src/types.rs:4-19 */
using Real = double;
using GridSize = std::array<std::size_t, 2>;
using CellPhysicalSize = std::array<Real, 2>;
using Velocity = std::array<Real, 2>;
using GridIndex = std::pair<std::size_t, std::size_t>;

/* GridArray<T> = ndarray::Array<T, Ix2> in its default (row-major) order: element (x, y) at
 * x * ny + y, y contiguous (src/types.rs:8-17). */
template <class T>
class GridArray {
  public:
    GridArray() : size_{0, 0} {}
    explicit GridArray(GridSize size, const T &fill = T()) : size_(size), data_(size[0] * size[1], fill) {}
    static GridArray zeros(GridSize size) { return GridArray(size, T()); }
    static GridArray from_elem(GridSize size, const T &elem) { return GridArray(size, elem); }

    GridSize dim() const { return size_; }
    std::size_t len() const { return data_.size(); }
    T *data() { return data_.data(); }
    const T *data() const { return data_.data(); }
    T &operator()(std::size_t x, std::size_t y) { return data_[x * size_[1] + y]; }
    const T &operator()(std::size_t x, std::size_t y) const { return data_[x * size_[1] + y]; }
    T &operator[](GridIndex i) { return (*this)(i.first, i.second); }
    const T &operator[](GridIndex i) const { return (*this)(i.first, i.second); }
    /* bounds-checked like ndarray's indexing (which panics) */
    T &at(std::size_t x, std::size_t y) {
        if (x >= size_[0] || y >= size_[1]) throw std::out_of_range("GridArray index");
        return (*this)(x, y);
    }
    const T &at(std::size_t x, std::size_t y) const { return const_cast<GridArray *>(this)->at(x, y); }
    bool operator==(const GridArray &o) const { return size_ == o.size_ && data_ == o.data_; }
    bool operator!=(const GridArray &o) const { return !(*this == o); }
    auto begin() { return data_.begin(); }
    auto end() { return data_.end(); }
    auto begin() const { return data_.begin(); }
    auto end() const { return data_.end(); }

  private:
    GridSize size_;
    std::vector<T> data_;
};

/* ---- cells (src/cell.rs:6-23) ----------------------------------------------------------- */

struct BoundaryCell {
    /* values are the C ABI's sb_kind codes */
    enum class Kind : std::uint8_t {
        NoSlip = SB_KIND_NOSLIP,
        Outflow = SB_KIND_OUTFLOW,
        Inflow = SB_KIND_INFLOW,
        MovingWall = SB_KIND_MOVING_WALL /* extension (lid-driven cavity), not in the reference */
    };
    Kind kind = Kind::NoSlip;
    Velocity velocity{0.0, 0.0}; /* payload of Inflow { velocity } (and MovingWall) */

    static BoundaryCell Inflow(Velocity velocity) { return {Kind::Inflow, velocity}; }
    static BoundaryCell Outflow() { return {Kind::Outflow, {0.0, 0.0}}; }
    static BoundaryCell NoSlip() { return {Kind::NoSlip, {0.0, 0.0}}; }
    static BoundaryCell MovingWall(Velocity velocity) { return {Kind::MovingWall, velocity}; }

    bool has_velocity() const { return kind == Kind::Inflow || kind == Kind::MovingWall; }
    bool operator==(const BoundaryCell &o) const {
        return kind == o.kind && (!has_velocity() || velocity == o.velocity);
    }
    bool operator!=(const BoundaryCell &o) const { return !(*this == o); }
    /* `{:?}` of the Rust enum, which is also its Display (src/cell.rs:12-16) */
    std::string to_string() const {
        auto num = [](Real r) {
            std::string s = std::to_string(r);
            while (s.size() > 1 && s.back() == '0' && s[s.size() - 2] != '.') s.pop_back();
            return s;
        };
        switch (kind) {
        case Kind::NoSlip: return "NoSlip";
        case Kind::Outflow: return "Outflow";
        case Kind::Inflow:
            return "Inflow { velocity: [" + num(velocity[0]) + ", " + num(velocity[1]) + "] }";
        default:
            return "MovingWall { velocity: [" + num(velocity[0]) + ", " + num(velocity[1]) + "] }";
        }
    }
};

class Cell {
  public:
    Cell() = default; /* Cell::Fluid */
    static Cell Fluid() { return Cell(); }
    static Cell Boundary(BoundaryCell b) {
        Cell c;
        c.boundary_ = b;
        return c;
    }
    bool is_fluid() const { return !boundary_.has_value(); }
    bool is_boundary() const { return boundary_.has_value(); }
    const BoundaryCell &boundary() const { return boundary_.value(); }
    std::uint8_t kind_code() const {
        return boundary_ ? static_cast<std::uint8_t>(boundary_->kind) : std::uint8_t(SB_KIND_FLUID);
    }
    bool operator==(const Cell &o) const { return boundary_ == o.boundary_; }
    bool operator!=(const Cell &o) const { return !(*this == o); }
    std::string to_string() const { /* src/cell.rs:25-29 */
        return boundary_ ? "Boundary(" + boundary_->to_string() + ")" : std::string("Fluid");
    }

  private:
    std::optional<BoundaryCell> boundary_;
};

/* ---- boundary classification (src/grid/mod.rs:19-49, 61-69) --------------------------- */

struct EdgeType {
    enum class Kind : std::uint8_t {
        North = SB_EDGE_N, NorthEast = SB_EDGE_NE, East = SB_EDGE_E, SouthEast = SB_EDGE_SE,
        South = SB_EDGE_S, SouthWest = SB_EDGE_SW, West = SB_EDGE_W, NorthWest = SB_EDGE_NW
    };
    Kind kind;
    /* the named neighbours of the Rust variants; "north" is y - 1 (src/grid/mod.rs:167-199) */
    std::optional<GridIndex> north_neighbor, east_neighbor, south_neighbor, west_neighbor;

    static EdgeType of(Kind kind, GridIndex cell) {
        EdgeType e{kind, {}, {}, {}, {}};
        const bool n = kind == Kind::North || kind == Kind::NorthEast || kind == Kind::NorthWest;
        const bool s = kind == Kind::South || kind == Kind::SouthEast || kind == Kind::SouthWest;
        const bool ea = kind == Kind::East || kind == Kind::NorthEast || kind == Kind::SouthEast;
        const bool w = kind == Kind::West || kind == Kind::NorthWest || kind == Kind::SouthWest;
        if (n) e.north_neighbor = GridIndex(cell.first, cell.second - 1);
        if (s) e.south_neighbor = GridIndex(cell.first, cell.second + 1);
        if (ea) e.east_neighbor = GridIndex(cell.first + 1, cell.second);
        if (w) e.west_neighbor = GridIndex(cell.first - 1, cell.second);
        return e;
    }
    bool operator==(const EdgeType &o) const {
        return kind == o.kind && north_neighbor == o.north_neighbor &&
               east_neighbor == o.east_neighbor && south_neighbor == o.south_neighbor &&
               west_neighbor == o.west_neighbor;
    }
    bool operator!=(const EdgeType &o) const { return !(*this == o); }
    const char *name() const {
        static const char *names[] = {"None", "North", "NorthEast", "East", "SouthEast",
                                      "South", "SouthWest", "West", "NorthWest"};
        return names[static_cast<int>(kind)];
    }
};

struct BoundaryList {
    std::vector<std::pair<GridIndex, std::optional<EdgeType>>> sorted_boundary_list;
    Real fluid_cells = 0.0;
};

/* ---- errors (src/simulation.rs:22-28, src/grid/mod.rs:51-59) --------------------------- */

class SimulationError : public std::runtime_error {
  public:
    SimulationError(sb_status status, const std::string &what) : std::runtime_error(what), status_(status) {}
    sb_status status() const { return status_; }

  private:
    sb_status status_;
};
/* SimulationGridError::BoundaryTooThinError(cell, index) */
class BoundaryTooThinError : public SimulationError {
  public:
    BoundaryTooThinError(GridIndex index, std::uint8_t kind, const std::string &what)
        : SimulationError(SB_BOUNDARY_TOO_THIN, what), index(index), kind(kind) {}
    GridIndex index;
    std::uint8_t kind; /* sb_kind of the offending cell */
};
/* SimulationGridError::BoundaryListIncorrectError(cell, index) */
class BoundaryListIncorrectError : public SimulationError {
  public:
    explicit BoundaryListIncorrectError(const std::string &what)
        : SimulationError(SB_BOUNDARY_LIST_INCORRECT, what) {}
};
class CudaError : public SimulationError { /* no GPU / driver failure: the path has no CPU fallback */
  public:
    explicit CudaError(const std::string &what) : SimulationError(SB_CUDA_ERROR, what) {}
};
class InvalidArgument : public SimulationError {
  public:
    explicit InvalidArgument(const std::string &what) : SimulationError(SB_INVALID_ARGUMENT, what) {}
};

namespace detail {
inline const char *kind_name(std::uint8_t kind) {
    switch (kind) {
    case SB_KIND_FLUID: return "Fluid";
    case SB_KIND_NOSLIP: return "Boundary(NoSlip)";
    case SB_KIND_OUTFLOW: return "Boundary(Outflow)";
    case SB_KIND_INFLOW: return "Boundary(Inflow)";
    default: return "Boundary(MovingWall)";
    }
}
/* sb_status -> the reference's error enum; `sim` may be null for a failed construction */
inline void check(sb_status st, const sb_sim *sim) {
    if (st == SB_OK) return;
    const char *m = sb_last_error_string();
    const std::string msg = m ? m : "";
    switch (st) {
    case SB_BOUNDARY_TOO_THIN: {
        std::uint64_t xy[2] = {0, 0};
        std::uint8_t kind = 0;
        sb_error_cell(sim, xy, &kind);
        /* "An error occurred with the SimulationGrid: `A cell `..` at `..` has fluid on
         * opposing sides.`" (src/simulation.rs:26-27, src/grid/mod.rs:57-58) */
        throw BoundaryTooThinError(
            GridIndex(xy[0], xy[1]), kind,
            std::string("BoundaryTooThinError: A cell `") + kind_name(kind) + "` at `(" +
                std::to_string(xy[0]) + ", " + std::to_string(xy[1]) +
                ")` has fluid on opposing sides.");
    }
    case SB_BOUNDARY_LIST_INCORRECT:
        throw BoundaryListIncorrectError("BoundaryListIncorrectError: " + msg);
    case SB_CUDA_ERROR: throw CudaError(msg);
    default: throw InvalidArgument(msg);
    }
}
} // namespace detail

/* ---- unfinalized (host-side) forms ------------------------------------------------------- */

/* UnfinalizedSimulationGrid (src/grid/mod.rs:85-92) */
struct UnfinalizedSimulationGrid {
    GridSize size{0, 0};
    GridArray<Real> pressure, u, v;
    GridArray<Cell> cell_type;
};

/* UnfinalizedSimulation (src/simulation.rs:30-44) */
struct UnfinalizedSimulation {
    GridSize size{0, 0};
    CellPhysicalSize cell_size{0.0, 0.0};
    Real delt = 0.0, gamma = 0.0, reynolds = 0.0;
    std::optional<Real> initial_norm_squared;
    Real sor_absolute_epsilon = 0.0;
    std::uint32_t max_iterations = 0, iterations = 0;
    Real time = 0.0, omega = 0.0;
    UnfinalizedSimulationGrid grid;
};

/* what the reference does not have: how the B200 build runs the solve */
struct Extensions {
    sb_sor_mode sor_mode = SB_SOR_REFERENCE_ORDER; /* bit-for-bit drop-in; SB_SOR_RED_BLACK = performance mode */
    int temporal_block = 0;                        /* red-black sweeps fused per pass (0 = default) */
    Real tau = 0.0;                                /* > 0: adaptive delt (NaSt2D COMP_delt) */
    int device = -1;                               /* CUDA ordinal, -1 = current */
};

class Simulation;

/* ---- SimulationGrid (src/grid/mod.rs:112-125), device-resident ------------------------- */

class SimulationGrid {
  public:
    GridSize size{0, 0};

    /* pub pressure / u / v / cell_type: download on read, upload on write */
    GridArray<Real> pressure() const { return get(SB_FIELD_P); }
    GridArray<Real> u() const { return get(SB_FIELD_U); }
    GridArray<Real> v() const { return get(SB_FIELD_V); }
    void set_pressure(const GridArray<Real> &a) { put(SB_FIELD_P, a); }
    void set_u(const GridArray<Real> &a) { put(SB_FIELD_U, a); }
    void set_v(const GridArray<Real> &a) { put(SB_FIELD_V, a); }

    GridArray<Cell> cell_type() const {
        std::vector<std::uint8_t> kind(size[0] * size[1]);
        detail::check(sb_download(h_, SB_FIELD_KIND, kind.data()), h_);
        std::size_t n = 0;
        detail::check(sb_get_boundary_velocities(h_, nullptr, 0, &n), h_);
        std::vector<sb_boundary_velocity> tab(n ? n : 1);
        detail::check(sb_get_boundary_velocities(h_, tab.data(), n, &n), h_);
        GridArray<Cell> out(size);
        for (std::size_t i = 0; i < kind.size(); ++i)
            if (kind[i] != SB_KIND_FLUID)
                out.data()[i] = Cell::Boundary({static_cast<BoundaryCell::Kind>(kind[i]), {0.0, 0.0}});
        for (std::size_t i = 0; i < n; ++i) {
            const std::size_t lin = tab[i].x * size[1] + tab[i].y;
            if (kind[lin] == SB_KIND_INFLOW || kind[lin] == SB_KIND_MOVING_WALL)
                out.data()[lin] = Cell::Boundary({static_cast<BoundaryCell::Kind>(kind[lin]), {tab[i].u, tab[i].v}});
        }
        return out;
    }
    /* writing `cell_type` (src/lib.rs:56-69 writes it, then calls rebuild_boundary_list at
     * :70): the caller rebuilds the list afterwards, exactly as there */
    void set_cell_type(const GridArray<Cell> &cells) {
        if (cells.dim() != size) throw InvalidArgument("cell_type: wrong shape");
        std::vector<std::uint8_t> kind;
        std::vector<sb_boundary_velocity> tab;
        flatten(cells, kind, tab);
        detail::check(sb_upload(h_, SB_FIELD_KIND, kind.data()), h_);
        detail::check(sb_set_boundary_velocities(h_, tab.data(), tab.size()), h_);
    }

    /* pub boundaries: BoundaryList (sorted_boundary_list in x-major order + fluid_cells) */
    BoundaryList boundaries() const {
        std::uint64_t n = 0;
        detail::check(sb_boundary_list(h_, nullptr, nullptr, 0, &n), h_);
        std::vector<std::uint64_t> idx(n ? n : 1);
        std::vector<std::uint8_t> edge(n ? n : 1);
        if (n) detail::check(sb_boundary_list(h_, idx.data(), edge.data(), n, &n), h_);
        BoundaryList bl;
        bl.sorted_boundary_list.reserve(n);
        for (std::uint64_t i = 0; i < n; ++i) {
            const GridIndex cell(idx[i] / size[1], idx[i] % size[1]);
            std::optional<EdgeType> e;
            if (edge[i] != SB_EDGE_NONE) e = EdgeType::of(static_cast<EdgeType::Kind>(edge[i]), cell);
            bl.sorted_boundary_list.emplace_back(cell, e);
        }
        bl.fluid_cells = state().fluid_cells;
        return bl;
    }
    std::array<Real, 2> pressure_range() const {
        const sb_state s = state();
        return {s.pressure_range[0], s.pressure_range[1]};
    }
    std::array<Real, 2> speed_range() const {
        const sb_state s = state();
        return {s.speed_range[0], s.speed_range[1]};
    }

    /* pub fns */
    void rebuild_boundary_list() { detail::check(sb_rebuild_boundary_list(h_), h_); }        /* :202-235 */
    void calculate_pressure_range() { detail::check(sb_calculate_pressure_range(h_), h_); }  /* :237-251 */
    void calculate_speed_range() { detail::check(sb_calculate_speed_range(h_), h_); }        /* :253-268 */
    void copy_pressure_to_boundaries() { detail::check(sb_copy_pressure_to_boundaries(h_), h_); } /* :343-412 */
    void set_boundary_u_and_v() { detail::check(sb_set_boundary_u_and_v(h_), h_); }          /* :414-651 */
    /* draw_cells (src/lib.rs:38-78): paint the 2x2 block at (m_x, m_y); false = rolled back
     * because the wall would be too thin */
    bool draw_cells(const Cell &cell, std::size_t m_x, std::size_t m_y) {
        std::int32_t applied = 0;
        const Velocity vel = cell.is_boundary() ? cell.boundary().velocity : Velocity{0.0, 0.0};
        detail::check(sb_edit_cells(h_, m_x, m_y, cell.kind_code(), vel[0], vel[1], &applied), h_);
        return applied != 0;
    }

    /* Cell grid -> u8 kinds + sparse velocity table of the C ABI */
    static void flatten(const GridArray<Cell> &cells, std::vector<std::uint8_t> &kind,
                        std::vector<sb_boundary_velocity> &tab) {
        const GridSize sz = cells.dim();
        kind.resize(sz[0] * sz[1]);
        tab.clear();
        for (std::size_t x = 0; x < sz[0]; ++x)
            for (std::size_t y = 0; y < sz[1]; ++y) {
                const Cell &c = cells(x, y);
                kind[x * sz[1] + y] = c.kind_code();
                if (c.is_boundary() && c.boundary().has_velocity())
                    tab.push_back({x, y, c.boundary().velocity[0], c.boundary().velocity[1]});
            }
    }

  private:
    friend class Simulation;
    sb_sim *h_ = nullptr;

    sb_state state() const {
        sb_state s{};
        detail::check(sb_get_state(h_, &s), h_);
        return s;
    }
    GridArray<Real> get(sb_field f) const {
        GridArray<Real> a(size);
        detail::check(sb_download(h_, f, a.data()), h_);
        return a;
    }
    void put(sb_field f, const GridArray<Real> &a) {
        if (a.dim() != size) throw InvalidArgument("field: wrong shape");
        detail::check(sb_upload(h_, f, a.data()), h_);
    }
};

/* ---- Simulation (src/simulation.rs:49-69) ---------------------------------------------- */

class Simulation {
  public:
    /* pub fields that never change after construction */
    GridSize size{0, 0};
    CellPhysicalSize cell_size{0.0, 0.0};
    /* pub grid: SimulationGrid */
    SimulationGrid grid;

    /* Simulation::try_from(UnfinalizedSimulation) (src/simulation.rs:71-99): classifies the
     * boundaries (BoundaryTooThinError), ranges, F/G, RHS, initial residual norm */
    static Simulation try_from(const UnfinalizedSimulation &item, const Extensions &ext = {}) {
        const GridSize sz = item.size;
        if (item.grid.size != sz || item.grid.cell_type.dim() != sz)
            throw InvalidArgument("UnfinalizedSimulation: grid size differs from size");
        for (const GridArray<Real> *a : {&item.grid.pressure, &item.grid.u, &item.grid.v})
            if (a->len() != 0 && a->dim() != sz) throw InvalidArgument("UnfinalizedSimulation: field shape");
        sb_params prm = params_of(item, ext);
        std::vector<std::uint8_t> kind;
        std::vector<sb_boundary_velocity> tab;
        SimulationGrid::flatten(item.grid.cell_type, kind, tab);
        auto ptr = [](const GridArray<Real> &a) { return a.len() ? a.data() : nullptr; };
        sb_sim *h = nullptr;
        detail::check(sb_create(&prm, ptr(item.grid.pressure), ptr(item.grid.u), ptr(item.grid.v),
                                kind.data(), tab.data(), tab.size(), &h),
                      nullptr);
        return Simulation(h, prm);
    }
    /* device-side mask generation (no host arrays): preset 1 simple_inflow, 2 obstacle, ...
     * (sb_create_preset); `grid` of `item` is ignored */
    static Simulation from_preset(const UnfinalizedSimulation &item, int preset,
                                  const std::vector<Real> &args = {}, const Extensions &ext = {}) {
        sb_params prm = params_of(item, ext);
        sb_sim *h = nullptr;
        detail::check(sb_create_preset(&prm, preset, args.data(), args.size(), &h), nullptr);
        return Simulation(h, prm);
    }

    Simulation(Simulation &&o) noexcept { *this = std::move(o); }
    Simulation &operator=(Simulation &&o) noexcept {
        if (this != &o) {
            reset();
            h_ = o.h_;
            prm_ = o.prm_;
            size = o.size;
            cell_size = o.cell_size;
            grid = o.grid;
            o.h_ = nullptr;
            o.grid.h_ = nullptr;
        }
        return *this;
    }
    Simulation(const Simulation &) = delete;
    Simulation &operator=(const Simulation &) = delete;
    ~Simulation() { reset(); }

    /* ---- the hot path: run_simulation_tick (src/simulation.rs:324-333) ---- */
    std::pair<std::uint32_t, Real> run_simulation_tick() {
        std::uint32_t it = 0;
        Real norm = 0.0;
        detail::check(sb_tick(h_, &it, &norm), h_);
        return {it, norm};
    }
    /* n ticks without reading anything back in between (src/lib.rs:214-219 ticks 20x per frame) */
    std::pair<std::uint32_t, Real> run_ticks(std::uint32_t n) {
        std::uint32_t it = 0;
        Real norm = 0.0;
        detail::check(sb_run_ticks(h_, n, &it, &norm), h_);
        return {it, norm};
    }

    /* one tick with HOST arrays authoritative (sb_tick_host): uploads p, u, v ([nx][ny] f64),
     * ticks, downloads them into the outputs (which may alias the inputs), one synchronisation;
     * include/stroemung_b200_pipeline.hpp keeps several such requests in flight */
    std::pair<std::uint32_t, Real> tick_host(const Real *p_in, const Real *u_in, const Real *v_in,
                                             Real *p_out, Real *u_out, Real *v_out) {
        std::uint32_t it = 0;
        Real norm = 0.0;
        detail::check(sb_tick_host(h_, p_in, u_in, v_in, p_out, u_out, v_out, &it, &norm), h_);
        return {it, norm};
    }

    /* stage functions of the reference, in tick order */
    void calculate_f_and_g() { detail::check(sb_calculate_f_and_g(h_), h_); } /* :122-202 */
    void calculate_rhs() { detail::check(sb_calculate_rhs(h_), h_); }         /* :204-214 */
    Real calculate_norm_squared() {                                           /* :216-227 */
        Real n = 0.0;
        detail::check(sb_calculate_norm_squared(h_, &n), h_);
        return n;
    }
    std::pair<std::uint32_t, Real> solve_sor() {                              /* :239-285 */
        std::uint32_t it = 0;
        Real norm = 0.0;
        detail::check(sb_solve_sor(h_, &it, &norm), h_);
        return {it, norm};
    }
    void set_u_and_v() { detail::check(sb_set_u_and_v(h_), h_); }             /* :287-322 */

    /* pub f / g / rhs (serde(skip) scratch arrays, src/simulation.rs:55-61) */
    GridArray<Real> f() const { return grid.get(SB_FIELD_F); }
    GridArray<Real> g() const { return grid.get(SB_FIELD_G); }
    GridArray<Real> rhs() const { return grid.get(SB_FIELD_RHS); }
    void set_f(const GridArray<Real> &a) { grid.put(SB_FIELD_F, a); }
    void set_g(const GridArray<Real> &a) { grid.put(SB_FIELD_G, a); }
    void set_rhs(const GridArray<Real> &a) { grid.put(SB_FIELD_RHS, a); }

    /* pub scalar fields */
    Real delt() const { return grid.state().delt; }
    Real gamma() const { return prm_.gamma; }
    Real reynolds() const { return prm_.reynolds; }
    Real sor_absolute_epsilon() const { return prm_.sor_absolute_epsilon; }
    std::uint32_t max_iterations() const { return prm_.max_iterations; }
    Real omega() const { return prm_.omega; }
    std::uint32_t iterations() const { return grid.state().iterations; }
    Real time() const { return grid.state().time; }
    std::optional<Real> initial_norm_squared() const {
        const sb_state s = grid.state();
        return s.has_initial_norm ? std::optional<Real>(s.initial_norm_squared) : std::nullopt;
    }
    void set_delt(Real v) { prm_.delt = v; push(true); }
    void set_gamma(Real v) { prm_.gamma = v; push(false); }
    void set_reynolds(Real v) { prm_.reynolds = v; push(false); }
    void set_sor_absolute_epsilon(Real v) { prm_.sor_absolute_epsilon = v; push(false); }
    void set_max_iterations(std::uint32_t v) { prm_.max_iterations = v; push(false); }
    void set_omega(Real v) { prm_.omega = v; push(false); }
    /* None makes the next solve_sor latch the norm after its first sweep (:229-237, :276) */
    void set_initial_norm_squared(std::optional<Real> v) {
        sync_from_device(false);
        prm_.has_initial_norm = v ? 1 : 0;
        prm_.initial_norm_squared = v ? *v : 0.0;
        detail::check(sb_set_params(h_, &prm_), h_);
    }
    void set_extensions(const Extensions &ext) {
        prm_.sor_mode = ext.sor_mode;
        prm_.temporal_block = ext.temporal_block;
        prm_.tau = ext.tau;
        push(false);
    }

    /* the values `#[derive(Serialize)]` would write (src/simulation.rs:49-69): the host form */
    UnfinalizedSimulation to_unfinalized() const {
        UnfinalizedSimulation u;
        u.size = size;
        u.cell_size = cell_size;
        u.delt = delt();
        u.gamma = gamma();
        u.reynolds = reynolds();
        u.initial_norm_squared = initial_norm_squared();
        u.sor_absolute_epsilon = sor_absolute_epsilon();
        u.max_iterations = max_iterations();
        u.iterations = iterations();
        u.time = time();
        u.omega = omega();
        u.grid = {size, grid.pressure(), grid.u(), grid.v(), grid.cell_type()};
        return u;
    }

    sb_sim *handle() const { return h_; } /* for the C ABI's instrumentation calls */

  private:
    sb_sim *h_ = nullptr;
    sb_params prm_{};

    Simulation(sb_sim *h, const sb_params &prm) : h_(h), prm_(prm) {
        size = {static_cast<std::size_t>(prm.nx), static_cast<std::size_t>(prm.ny)};
        cell_size = {prm.delx, prm.dely};
        grid.size = size;
        grid.h_ = h;
    }
    void reset() {
        if (h_) sb_destroy(h_);
        h_ = nullptr;
        grid.h_ = nullptr;
    }
    static sb_params params_of(const UnfinalizedSimulation &item, const Extensions &ext) {
        sb_params p{};
        p.nx = item.size[0];
        p.ny = item.size[1];
        p.delx = item.cell_size[0];
        p.dely = item.cell_size[1];
        p.delt = item.delt;
        p.gamma = item.gamma;
        p.reynolds = item.reynolds;
        p.sor_absolute_epsilon = item.sor_absolute_epsilon;
        p.omega = item.omega;
        p.time = item.time;
        p.max_iterations = item.max_iterations;
        p.iterations = item.iterations;
        p.has_initial_norm = item.initial_norm_squared ? 1 : 0;
        p.initial_norm_squared = item.initial_norm_squared.value_or(0.0);
        p.sor_mode = ext.sor_mode;
        p.tau = ext.tau;
        p.temporal_block = ext.temporal_block;
        p.device = ext.device;
        return p;
    }
    /* sb_set_params takes the whole scalar set: refresh the device-owned ones first */
    void sync_from_device(bool keep_delt) {
        const sb_state s = grid.state();
        prm_.time = s.time;
        prm_.iterations = s.iterations;
        prm_.has_initial_norm = s.has_initial_norm;
        prm_.initial_norm_squared = s.initial_norm_squared;
        if (prm_.tau > 0 && !keep_delt) prm_.delt = s.delt;
    }
    void push(bool keep_delt) {
        sync_from_device(keep_delt);
        detail::check(sb_set_params(h_, &prm_), h_);
    }
};

/* ---- presets (src/grid/presets.rs:8-87): the host-side masks, as UnfinalizedSimulationGrid
 * (the reference returns a finalized SimulationGrid and converts it back with `.into()`,
 * src/simulation.rs:443, :587) ------------------------------------------------------------ */
namespace presets {

inline UnfinalizedSimulationGrid empty(GridSize size) { /* :8-17 */
    return {size, GridArray<Real>::zeros(size), GridArray<Real>::zeros(size),
            GridArray<Real>::zeros(size), GridArray<Cell>::from_elem(size, Cell::Fluid())};
}

inline UnfinalizedSimulationGrid simple_inflow(GridSize size) { /* :19-40 */
    UnfinalizedSimulationGrid g = empty(size);
    for (std::size_t x = 0; x < size[0]; ++x) {
        g.cell_type(x, 0) = Cell::Boundary(BoundaryCell::NoSlip());
        g.cell_type(x, size[1] - 1) = Cell::Boundary(BoundaryCell::NoSlip());
    }
    for (std::size_t y = 1; y + 1 < size[1]; ++y) {
        g.cell_type(0, y) = Cell::Boundary(BoundaryCell::Inflow({1.0, 0.0}));
        g.cell_type(size[0] - 1, y) = Cell::Boundary(BoundaryCell::Outflow());
    }
    return g;
}

/* draw_circle (:42-62): integer rasteriser, `radius as usize` truncation, saturating bounds */
inline void draw_circle(GridArray<Cell> &cell_array, std::size_t x, std::size_t y, Real radius) {
    const GridSize sz = cell_array.dim();
    const std::size_t r = static_cast<std::size_t>(radius);
    for (std::size_t xi = x >= r ? x - r : 0; xi < x + r; ++xi) {
        if (xi >= sz[0]) continue;
        const std::int32_t x_dist = static_cast<std::int32_t>(xi) - static_cast<std::int32_t>(x);
        for (std::size_t yi = y >= r ? y - r : 0; yi < y + r; ++yi) {
            if (yi >= sz[1]) continue;
            const std::int32_t y_dist = static_cast<std::int32_t>(yi) - static_cast<std::int32_t>(y);
            const Real distance = __builtin_sqrt(static_cast<Real>(x_dist * x_dist + y_dist * y_dist));
            if (distance < radius) cell_array(xi, yi) = Cell::Boundary(BoundaryCell::NoSlip());
        }
    }
}

inline UnfinalizedSimulationGrid obstacle(GridSize size) { /* :64-87 */
    UnfinalizedSimulationGrid g = simple_inflow(size);
    draw_circle(g.cell_type, 20, size[1] / 2, 5.0);
    return g;
}

} // namespace presets

/* ---- cell-level operators (src/math.rs:19-186, src/simulation.rs:349-392) ---------------
 * 3x3 views in the reference's [x][y] order; evaluated ON THE DEVICE by the same __device__
 * functions the kernels use, so the reference's known-answer tests run against the CUDA path. */
using View3x3 = std::array<std::array<Real, 3>, 3>;

namespace math {
inline Real du2dx(const View3x3 &u, Real delx, Real gamma) { /* :19 */
    Real o = 0.0;
    detail::check(sb_du2dx(&u[0][0], delx, gamma, &o), nullptr);
    return o;
}
inline Real duvdx(const View3x3 &u, const View3x3 &v, Real delx, Real gamma) { /* :53 */
    Real o = 0.0;
    detail::check(sb_duvdx(&u[0][0], &v[0][0], delx, gamma, &o), nullptr);
    return o;
}
inline Real duvdy(const View3x3 &u, const View3x3 &v, Real dely, Real gamma) { /* :97 */
    Real o = 0.0;
    detail::check(sb_duvdy(&u[0][0], &v[0][0], dely, gamma, &o), nullptr);
    return o;
}
inline Real dv2dy(const View3x3 &v, Real dely, Real gamma) { /* :136 */
    Real o = 0.0;
    detail::check(sb_dv2dy(&v[0][0], dely, gamma, &o), nullptr);
    return o;
}
inline Real laplacian(const View3x3 &e, Real delx, Real dely) { /* :162 */
    Real o = 0.0;
    detail::check(sb_laplacian(&e[0][0], delx, dely, &o), nullptr);
    return o;
}
inline Real residual(const View3x3 &p, Real delx, Real dely, Real rhs) { /* :176 */
    Real o = 0.0;
    detail::check(sb_residual(&p[0][0], delx, dely, rhs, &o), nullptr);
    return o;
}
} // namespace math

inline Real calculate_f(const View3x3 &u, const View3x3 &v, Real delx, Real dely, Real delt,
                        Real gamma, Real reynolds) { /* src/simulation.rs:349-363 */
    Real o = 0.0;
    detail::check(sb_calculate_f(&u[0][0], &v[0][0], delx, dely, delt, gamma, reynolds, &o), nullptr);
    return o;
}
inline Real calculate_g(const View3x3 &u, const View3x3 &v, Real delx, Real dely, Real delt,
                        Real gamma, Real reynolds) { /* src/simulation.rs:378-392 */
    Real o = 0.0;
    detail::check(sb_calculate_g(&u[0][0], &v[0][0], delx, dely, delt, gamma, reynolds, &o), nullptr);
    return o;
}

} // namespace stroemung

#endif /* STROEMUNG_B200_HPP */
