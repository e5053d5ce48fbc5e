// The reference's own unit tests, replayed in C++ through include/stroemung_b200.hpp (the
// host-side mirror of the Rust API) against libstroemung_b200.so on a B200.
//
// Each test is named after the Rust test it restates and cites it.  Expected values are the
// reference-held golden vectors of tests/golden/*.json (golden.inc is generated from them by
// make_golden_inc.py); where the reference has no vector (grids beyond 4x3) the CPU oracle
// (oracle/stroemung_oracle.h, test infrastructure) is the checker.  Fields are compared on bit
// patterns, like `assert_eq!` and the insta snapshots; residual norms within 1e-12 relative.
//
//   test_reference_tests            run everything (needs a GPU; exit code = failed tests)
//   test_reference_tests --list     print the test names, touch nothing (CPU check)
//   test_reference_tests --host     run the host_* tests only (file format; no device needed)
//   test_reference_tests NAME...    run the named tests
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <memory>
#include <sstream>
#include <string>
#include <tuple>
#include <vector>

#include "stroemung_b200.hpp"
#include "stroemung_b200_json.hpp"
#include "stroemung_b200_pipeline.hpp"
#include "stroemung_oracle.h"

#include "golden.inc"

using namespace stroemung;

// ---- a very small test harness --------------------------------------------------------
struct Failure {
    std::string what;
};
#define REQUIRE(cond)                                                                        \
    do {                                                                                     \
        if (!(cond))                                                                         \
            throw Failure{std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": " #cond}; \
    } while (0)

static std::vector<std::pair<std::string, std::function<void()>>> &registry() {
    static std::vector<std::pair<std::string, std::function<void()>>> r;
    return r;
}
struct Registrar {
    Registrar(const char *name, std::function<void()> fn) { registry().emplace_back(name, fn); }
};
#define TEST(name)                                   \
    static void name();                              \
    static Registrar reg_##name(#name, name);        \
    static void name()

static bool same_bits(double a, double b) { return std::memcmp(&a, &b, sizeof a) == 0; }
// residual norms are sums: the kernels add the reference's terms in a tree, so norms are held to
// 1e-12 relative ("summation order beyond 2 terms is parity-unpinned")
static bool close(double a, double b) {
    return a == b || std::abs(a - b) <= 1e-12 * std::max(std::abs(a), std::abs(b));
}
static void require_bits(const GridArray<Real> &got, const double *want, const char *what) {
    for (std::size_t i = 0; i < got.len(); ++i)
        if (!same_bits(got.data()[i], want[i])) {
            char buf[256];
            std::snprintf(buf, sizeof buf, "%s[%zu]: got %a, want %a", what, i, got.data()[i], want[i]);
            throw Failure{buf};
        }
}
static View3x3 view(const double m[3][3]) {
    View3x3 v;
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) v[a][b] = m[a][b];
    return v;
}

// ---- src/math.rs:192-402 (24 exact known-answer tests), evaluated on the device ---------
TEST(math_test_du2dx) { // src/math.rs:197-240
    for (const Kat1 &c : KAT_DU2DX) REQUIRE(same_bits(math::du2dx(view(c.a), c.d, c.gamma), c.expected));
}
TEST(math_test_dv2dy) { // src/math.rs:242-285
    for (const Kat1 &c : KAT_DV2DY) REQUIRE(same_bits(math::dv2dy(view(c.a), c.d, c.gamma), c.expected));
}
TEST(math_test_duvdx) { // src/math.rs:287-324
    for (const Kat2 &c : KAT_DUVDX)
        REQUIRE(same_bits(math::duvdx(view(c.u), view(c.v), c.d, c.gamma), c.expected));
}
TEST(math_test_duvdy) { // src/math.rs:326-363
    for (const Kat2 &c : KAT_DUVDY)
        REQUIRE(same_bits(math::duvdy(view(c.u), view(c.v), c.d, c.gamma), c.expected));
}
TEST(math_test_laplacian) { // src/math.rs:365-402
    for (const KatLap &c : KAT_LAPLACIAN)
        REQUIRE(same_bits(math::laplacian(view(c.e), c.delx, c.dely), c.expected));
}
// ---- src/simulation.rs:449-569 (8 KATs) ------------------------------------------------------
TEST(simulation_test_calculate_f) {
    for (const KatFG &c : KAT_CALCULATE_F)
        REQUIRE(same_bits(calculate_f(view(c.u), view(c.v), c.delx, c.dely, c.delt, c.gamma, c.reynolds),
                          c.expected));
}
TEST(simulation_test_calculate_g) {
    for (const KatFG &c : KAT_CALCULATE_G)
        REQUIRE(same_bits(calculate_g(view(c.u), view(c.v), c.delx, c.dely, c.delt, c.gamma, c.reynolds),
                          c.expected));
}

// ---- src/simulation.rs:571-618 `simulation_tick` ----------------------------------------------
static UnfinalizedSimulation tick_case(GridSize size, UnfinalizedSimulationGrid grid) {
    UnfinalizedSimulation u;
    u.size = size;
    u.cell_size = {0.1, 0.2};
    u.delt = 0.005;
    u.gamma = 0.9;
    u.reynolds = 100.0;
    u.sor_absolute_epsilon = 0.001;
    u.max_iterations = 100;
    u.initial_norm_squared = std::nullopt;
    u.iterations = 0;
    u.time = 0.0;
    u.omega = 1.7;
    u.grid = std::move(grid);
    return u;
}
static void require_sim(const Simulation &sim, const SimSnap &want, const char *what) {
    REQUIRE(sim.iterations() == want.iterations);
    REQUIRE(same_bits(sim.time(), want.time));
    REQUIRE(sim.initial_norm_squared().has_value());
    REQUIRE(same_bits(*sim.initial_norm_squared(), want.initial_norm_squared));
    require_bits(sim.grid.pressure(), want.p, what);
    require_bits(sim.grid.u(), want.u, what);
    require_bits(sim.grid.v(), want.v, what);
}
TEST(simulation_simulation_tick) {
    const GridSize size{4, 3};
    Simulation sim = Simulation::try_from(tick_case(size, presets::simple_inflow(size)));

    auto [sor_iterations, norm_squared] = sim.run_simulation_tick();
    require_bits(sim.f(), TICK_F_1, "f after tick 1");
    require_bits(sim.g(), TICK_G_1, "g after tick 1");
    require_bits(sim.rhs(), TICK_RHS_1, "rhs after tick 1");
    require_sim(sim, TICK_SIM_1, "sim after tick 1");
    REQUIRE(sor_iterations == TICK_ASSERT_ITERS[0]);          // assert_eq!(sor_iterations, 100)
    REQUIRE(close(norm_squared, TICK_ASSERT_NORM[0]));        // 562901.7447199143

    std::uint32_t last_sor_iterations = 0;
    Real last_norm_squared = 0.0;
    for (int i = 0; i < 100; ++i) std::tie(last_sor_iterations, last_norm_squared) = sim.run_simulation_tick();
    REQUIRE(last_sor_iterations == TICK_ASSERT_ITERS[1]);     // 1
    REQUIRE(close(last_norm_squared, TICK_ASSERT_NORM[1]));   // 3.8344148218167323e-20
    require_bits(sim.f(), TICK_F_101, "f after tick 101");
    require_bits(sim.g(), TICK_G_101, "g after tick 101");
    require_bits(sim.rhs(), TICK_RHS_101, "rhs after tick 101");
    require_sim(sim, TICK_SIM_101, "sim after tick 101");

    for (int i = 0; i < 100; ++i) sim.run_simulation_tick();
    require_sim(sim, TICK_SIM_201, "sim after tick 201");
}

// ---- src/simulation.rs:422-446 `serialize` -----------------------------------------------------
TEST(simulation_serialize) {
    const GridSize size{5, 7};
    UnfinalizedSimulation u = tick_case(size, presets::empty(size));
    u.cell_size = {1., 2.};
    u.delt = 1.4;
    u.gamma = 1.7;
    Simulation simulation = Simulation::try_from(u);
    const UnfinalizedSimulation out = simulation.to_unfinalized(); // what Serialize writes
    REQUIRE(out.size == size);
    REQUIRE(same_bits(out.delt, SERIALIZE_DELT));
    REQUIRE(out.initial_norm_squared.has_value());
    REQUIRE(same_bits(*out.initial_norm_squared, SERIALIZE_INITIAL_NORM));
    REQUIRE(out.iterations == 0 && out.max_iterations == 100);
    for (Real x : out.grid.pressure) REQUIRE(same_bits(x, 0.0));
    for (Real x : out.grid.u) REQUIRE(same_bits(x, 0.0));
    for (Real x : out.grid.v) REQUIRE(same_bits(x, 0.0));
    for (const Cell &c : out.grid.cell_type) REQUIRE(c == Cell::Fluid());
    REQUIRE(simulation.grid.boundaries().sorted_boundary_list.empty());
    REQUIRE(simulation.grid.boundaries().fluid_cells == 35.0);
}

// ---- src/grid/mod.rs:683-706 `thin_boundary` ---------------------------------------------------
static UnfinalizedSimulation grid_only(GridSize size, const std::vector<GridIndex> &noslip) {
    UnfinalizedSimulationGrid g = presets::empty(size);
    for (GridIndex idx : noslip) g.cell_type[idx] = Cell::Boundary(BoundaryCell::NoSlip());
    UnfinalizedSimulation u = tick_case(size, std::move(g));
    u.cell_size = {1., 1.};
    return u;
}
TEST(grid_thin_boundary) {
    const GridSize size{3, 3};
    const std::vector<std::vector<GridIndex>> boundaries = {{{1, 0}, {1, 1}, {1, 2}},
                                                            {{0, 1}, {1, 1}, {2, 1}}};
    for (const auto &example : boundaries) {
        bool is_err = false;
        try {
            Simulation::try_from(grid_only(size, example));
        } catch (const BoundaryTooThinError &e) {
            is_err = true;
            REQUIRE(std::string(e.what()).find("BoundaryTooThinError") != std::string::npos);
            // the first offender in x-major order is the one reported (:225-232)
            REQUIRE(e.index == example[0]);
            REQUIRE(e.kind == SB_KIND_NOSLIP);
        }
        REQUIRE(is_err);
    }
}

// ---- src/grid/mod.rs:708-801 `rebuild_boundary_list` -------------------------------------------
TEST(grid_rebuild_boundary_list) {
    using K = EdgeType::Kind;
    const GridSize size{3, 3};
    struct Example {
        std::vector<GridIndex> boundaries;
        std::vector<std::optional<EdgeType>> neighbors;
    };
    auto some = [](K k, GridIndex cell) { return std::optional<EdgeType>(EdgeType::of(k, cell)); };
    const std::vector<Example> examples = {
        // Everything except for the middle cell is a boundary
        {{{0, 0}, {0, 1}, {0, 2}, {1, 0}, {1, 2}, {2, 0}, {2, 1}, {2, 2}},
         {std::nullopt, some(K::East, {0, 1}), std::nullopt, some(K::South, {1, 0}),
          some(K::North, {1, 2}), std::nullopt, some(K::West, {2, 1}), std::nullopt}},
        // All corners are boundaries
        {{{0, 0}, {0, 2}, {2, 0}, {2, 2}},
         {some(K::SouthEast, {0, 0}), some(K::NorthEast, {0, 2}), some(K::SouthWest, {2, 0}),
          some(K::NorthWest, {2, 2})}},
    };
    // the neighbour indices the reference spells out (:727-764)
    REQUIRE(examples[0].neighbors[1]->east_neighbor == GridIndex(1, 1));
    REQUIRE(examples[0].neighbors[3]->south_neighbor == GridIndex(1, 1));
    REQUIRE(examples[0].neighbors[4]->north_neighbor == GridIndex(1, 1));
    REQUIRE(examples[0].neighbors[6]->west_neighbor == GridIndex(1, 1));
    REQUIRE(examples[1].neighbors[0]->south_neighbor == GridIndex(0, 1));
    REQUIRE(examples[1].neighbors[0]->east_neighbor == GridIndex(1, 0));
    REQUIRE(examples[1].neighbors[3]->north_neighbor == GridIndex(2, 1));
    REQUIRE(examples[1].neighbors[3]->west_neighbor == GridIndex(1, 2));

    for (const Example &ex : examples) {
        Simulation sim = Simulation::try_from(grid_only(size, ex.boundaries));
        const BoundaryList bl = sim.grid.boundaries();
        REQUIRE(bl.sorted_boundary_list.size() == ex.boundaries.size());
        for (std::size_t i = 0; i < ex.boundaries.size(); ++i) {
            REQUIRE(bl.sorted_boundary_list[i].first == ex.boundaries[i]);
            REQUIRE(bl.sorted_boundary_list[i].second == ex.neighbors[i]);
        }
        REQUIRE(bl.fluid_cells == Real(9 - ex.boundaries.size()));
    }
}

// ---- src/lib.rs:38-78 `draw_cells` + rebuild: edit, thin-wall roll-back -------------------------
TEST(lib_draw_cells_rolls_back_thin_walls) {
    const GridSize size{12, 10};
    Simulation sim = Simulation::try_from(tick_case(size, presets::simple_inflow(size)));
    const GridArray<Cell> before = sim.grid.cell_type();
    REQUIRE(before(0, 3) == Cell::Boundary(BoundaryCell::Inflow({1.0, 0.0})));
    // a 2x2 block in the open channel is a legal obstacle
    REQUIRE(sim.grid.draw_cells(Cell::Boundary(BoundaryCell::NoSlip()), 5, 4));
    GridArray<Cell> after = sim.grid.cell_type();
    REQUIRE(after(5, 4) == Cell::Boundary(BoundaryCell::NoSlip()) && after(6, 5) == after(5, 4));
    // painting fluid over one of its rows leaves a 1-cell wall with fluid on opposing
    // sides: rebuild_boundary_list fails and the edit is rolled back
    REQUIRE(!sim.grid.draw_cells(Cell::Fluid(), 5, 5));
    REQUIRE(sim.grid.cell_type() == after);
    sim.run_simulation_tick();
}

// ---- beyond the reference's fixtures: the default obstacle preset against the oracle ----------
static void against_oracle(sb_sor_mode mode, int temporal_block, GridSize size, int ticks) {
    UnfinalizedSimulation u = tick_case(size, presets::obstacle(size));
    Extensions ext;
    ext.sor_mode = mode;
    ext.temporal_block = temporal_block;
    Simulation sim = Simulation::try_from(u, ext);

    std::vector<std::uint8_t> kind;
    std::vector<sb_boundary_velocity> tab;
    SimulationGrid::flatten(u.grid.cell_type, kind, tab);
    std::vector<double> bu(kind.size(), 0.0), bv(kind.size(), 0.0);
    for (const sb_boundary_velocity &t : tab) {
        bu[t.x * size[1] + t.y] = t.u;
        bv[t.x * size[1] + t.y] = t.v;
    }
    so_params op{};
    op.nx = size[0];
    op.ny = size[1];
    op.delx = u.cell_size[0];
    op.dely = u.cell_size[1];
    op.delt = u.delt;
    op.gamma = u.gamma;
    op.reynolds = u.reynolds;
    op.sor_absolute_epsilon = u.sor_absolute_epsilon;
    op.omega = u.omega;
    op.max_iterations = u.max_iterations;
    op.sor_mode = mode == SB_SOR_RED_BLACK ? SO_SOR_RED_BLACK : SO_SOR_REFERENCE_ORDER;
    so_sim *ref = nullptr;
    std::uint64_t err[2];
    REQUIRE(so_create(&op, nullptr, nullptr, nullptr, kind.data(), bu.data(), bv.data(), &ref, err) == SO_OK);
    {
        so_state st0;
        so_get_state(ref, &st0);
        REQUIRE(close(*sim.initial_norm_squared(), st0.initial_norm_squared));
    }
    for (int t = 0; t < ticks; ++t) {
        std::uint32_t it = 0;
        double norm = 0.0;
        REQUIRE(so_tick(ref, &it, &norm) == SO_OK);
        auto [git, gnorm] = sim.run_simulation_tick();
        REQUIRE(git == it);
        REQUIRE(close(gnorm, norm));
        require_bits(sim.grid.pressure(), so_p(ref), "pressure");
        require_bits(sim.grid.u(), so_u(ref), "u");
        require_bits(sim.grid.v(), so_v(ref), "v");
        require_bits(sim.f(), so_f(ref), "f");
        require_bits(sim.g(), so_g(ref), "g");
        require_bits(sim.rhs(), so_rhs(ref), "rhs");
    }
    so_state st;
    so_get_state(ref, &st);
    REQUIRE(same_bits(sim.grid.speed_range()[0], st.speed_range[0]));
    REQUIRE(same_bits(sim.grid.speed_range()[1], st.speed_range[1]));
    REQUIRE(same_bits(sim.grid.pressure_range()[0], st.pressure_range[0]));
    REQUIRE(same_bits(sim.grid.pressure_range()[1], st.pressure_range[1]));
    REQUIRE(sim.grid.boundaries().fluid_cells == st.fluid_cells);
    so_destroy(ref);
}
TEST(obstacle_preset_reference_order_vs_oracle) { against_oracle(SB_SOR_REFERENCE_ORDER, 0, {100, 20}, 5); }
TEST(obstacle_preset_red_black_vs_oracle) { against_oracle(SB_SOR_RED_BLACK, 0, {100, 20}, 5); }
TEST(obstacle_channel_red_black_pass_kernels_vs_oracle) { against_oracle(SB_SOR_RED_BLACK, 4, {640, 300}, 2); }

// ---- the stage functions of the reference called one by one through the mirror --------------
// (`run_simulation_tick` is exactly this sequence, src/simulation.rs:324-333)
struct OracleTwin {
    so_sim *ref = nullptr;
    so_params op{};
    OracleTwin(const UnfinalizedSimulation &u, sb_sor_mode mode) {
        std::vector<std::uint8_t> kind;
        std::vector<sb_boundary_velocity> tab;
        SimulationGrid::flatten(u.grid.cell_type, kind, tab);
        std::vector<double> bu(kind.size(), 0.0), bv(kind.size(), 0.0);
        for (const sb_boundary_velocity &t : tab) {
            bu[t.x * u.size[1] + t.y] = t.u;
            bv[t.x * u.size[1] + t.y] = t.v;
        }
        op.nx = u.size[0];
        op.ny = u.size[1];
        op.delx = u.cell_size[0];
        op.dely = u.cell_size[1];
        op.delt = u.delt;
        op.gamma = u.gamma;
        op.reynolds = u.reynolds;
        op.sor_absolute_epsilon = u.sor_absolute_epsilon;
        op.omega = u.omega;
        op.max_iterations = u.max_iterations;
        op.sor_mode = mode == SB_SOR_RED_BLACK ? SO_SOR_RED_BLACK : SO_SOR_REFERENCE_ORDER;
        std::uint64_t err[2];
        if (so_create(&op, nullptr, nullptr, nullptr, kind.data(), bu.data(), bv.data(), &ref, err) != SO_OK)
            throw Failure{"oracle: so_create failed"};
    }
    ~OracleTwin() { so_destroy(ref); }
};
static void require_fields(const Simulation &sim, so_sim *ref, const char *when) {
    const std::string w(when);
    require_bits(sim.grid.pressure(), so_p(ref), (w + ": pressure").c_str());
    require_bits(sim.grid.u(), so_u(ref), (w + ": u").c_str());
    require_bits(sim.grid.v(), so_v(ref), (w + ": v").c_str());
    require_bits(sim.f(), so_f(ref), (w + ": f").c_str());
    require_bits(sim.g(), so_g(ref), (w + ": g").c_str());
    require_bits(sim.rhs(), so_rhs(ref), (w + ": rhs").c_str());
}
TEST(stage_functions_one_by_one_vs_oracle) {
    const GridSize size{100, 20};
    const UnfinalizedSimulation u = tick_case(size, presets::obstacle(size));
    Simulation sim = Simulation::try_from(u);
    OracleTwin o(u, SB_SOR_REFERENCE_ORDER);
    require_fields(sim, o.ref, "after try_from");
    for (int t = 0; t < 2; ++t) {
        sim.grid.set_boundary_u_and_v();
        REQUIRE(so_set_boundary_u_and_v(o.ref) == SO_OK);
        require_fields(sim, o.ref, "after set_boundary_u_and_v");
        sim.calculate_f_and_g();
        so_calculate_f_and_g(o.ref);
        sim.calculate_rhs();
        so_calculate_rhs(o.ref);
        require_fields(sim, o.ref, "after calculate_rhs");
        REQUIRE(close(sim.calculate_norm_squared(), so_calculate_norm_squared(o.ref)));
        std::uint32_t oit = 0;
        double onorm = 0.0;
        REQUIRE(so_solve_sor(o.ref, &oit, &onorm) == SO_OK);
        auto [it, norm] = sim.solve_sor();
        REQUIRE(it == oit);
        REQUIRE(close(norm, onorm));
        require_fields(sim, o.ref, "after solve_sor");
        sim.set_u_and_v();
        so_set_u_and_v(o.ref);
        require_fields(sim, o.ref, "after set_u_and_v");
        so_state st;
        so_get_state(o.ref, &st);
        REQUIRE(same_bits(sim.grid.speed_range()[1], st.speed_range[1]));
        REQUIRE(same_bits(sim.grid.pressure_range()[0], st.pressure_range[0]));
        REQUIRE(same_bits(sim.grid.pressure_range()[1], st.pressure_range[1]));
    }
    sim.grid.copy_pressure_to_boundaries();
    REQUIRE(so_copy_pressure_to_boundaries(o.ref) == SO_OK);
    require_fields(sim, o.ref, "after copy_pressure_to_boundaries");
}

// ---- pub scalar fields are writable (src/simulation.rs:50-69): the setters of the mirror -----
TEST(pub_fields_written_through_the_mirror) {
    const GridSize size{100, 20};
    const UnfinalizedSimulation u = tick_case(size, presets::obstacle(size));
    Simulation sim = Simulation::try_from(u);
    OracleTwin o(u, SB_SOR_REFERENCE_ORDER);
    sim.set_max_iterations(7);
    sim.set_omega(1.5);
    sim.set_delt(0.004);
    o.op.max_iterations = 7;
    o.op.omega = 1.5;
    o.op.delt = 0.004;
    so_state st;
    so_get_state(o.ref, &st);
    o.op.has_initial_norm = st.has_initial_norm;
    o.op.initial_norm_squared = st.initial_norm_squared;
    so_set_params(o.ref, &o.op);
    REQUIRE(sim.max_iterations() == 7 && sim.omega() == 1.5 && sim.delt() == 0.004);
    std::uint32_t oit = 0;
    double onorm = 0.0;
    REQUIRE(so_tick(o.ref, &oit, &onorm) == SO_OK);
    auto [it, norm] = sim.run_simulation_tick();
    REQUIRE(it == 7 && oit == 7);   // the obstacle preset always runs into the cap
    REQUIRE(close(norm, onorm));
    require_fields(sim, o.ref, "after a tick with edited parameters");
    REQUIRE(same_bits(sim.time(), 0.004));
    // initial_norm_squared = None: solve_sor latches the norm after its first sweep (:229-237)
    sim.set_initial_norm_squared(std::nullopt);
    so_clear_initial_norm(o.ref);
    REQUIRE(!sim.initial_norm_squared().has_value());
    REQUIRE(so_tick(o.ref, &oit, &onorm) == SO_OK);
    std::tie(it, norm) = sim.run_simulation_tick();
    REQUIRE(it == oit && close(norm, onorm));
    so_get_state(o.ref, &st);
    REQUIRE(sim.initial_norm_squared().has_value());
    REQUIRE(close(*sim.initial_norm_squared(), st.initial_norm_squared));
    require_fields(sim, o.ref, "after the latching tick");
    REQUIRE(sim.iterations() == 2);
}

// ---- device-side mask generation equals the host-side preset (src/grid/presets.rs:64-87) -----
TEST(device_preset_equals_host_preset) {
    const GridSize size{100, 20};
    const UnfinalizedSimulation u = tick_case(size, presets::obstacle(size));
    Simulation dev = Simulation::from_preset(u, /*obstacle*/ 2);
    Simulation host = Simulation::try_from(u);
    REQUIRE(dev.grid.cell_type() == host.grid.cell_type());
    REQUIRE(dev.grid.cell_type() == u.grid.cell_type);
    const BoundaryList a = dev.grid.boundaries(), b = host.grid.boundaries();
    REQUIRE(a.fluid_cells == b.fluid_cells);
    REQUIRE(a.sorted_boundary_list.size() == b.sorted_boundary_list.size());
    for (std::size_t i = 0; i < a.sorted_boundary_list.size(); ++i)
        REQUIRE(a.sorted_boundary_list[i] == b.sorted_boundary_list[i]);
    auto [it_d, norm_d] = dev.run_ticks(3);      // src/lib.rs:214-219 ticks in a row
    std::uint32_t it_h = 0;
    Real norm_h = 0.0;
    for (int t = 0; t < 3; ++t) std::tie(it_h, norm_h) = host.run_simulation_tick();
    REQUIRE(it_d == it_h && close(norm_d, norm_h));
    require_bits(dev.grid.pressure(), host.grid.pressure().data(), "pressure");
    require_bits(dev.grid.u(), host.grid.u().data(), "u");
    REQUIRE(dev.iterations() == 3 && host.iterations() == 3);
}

// ---- the reference's file format (include/stroemung_b200_json.hpp); host_* tests need no GPU --
static void require_parsed(const UnfinalizedSimulationGrid &g, const ParsedSnap &want, const char *what) {
    REQUIRE(g.size == (GridSize{want.nx, want.ny}));
    require_bits(g.pressure, want.p, what);
    require_bits(g.u, want.u, what);
    require_bits(g.v, want.v, what);
    for (std::size_t i = 0; i < g.cell_type.len(); ++i) {
        const Cell &c = g.cell_type.data()[i];
        REQUIRE(c.kind_code() == want.kind[i]);
        if (c.is_boundary() && c.boundary().has_velocity()) {
            REQUIRE(same_bits(c.boundary().velocity[0], want.bu[i]));
            REQUIRE(same_bits(c.boundary().velocity[1], want.bv[i]));
        }
    }
}
TEST(host_serde_json_number_parser) {
    // the literal of src/test_data/small_simulation_with_boundaries.json that serde_json 1.0.140
    // (no float_roundtrip) reads one ulp low; value from the reference's own snapshot
    const double got = json::serde_json_f64("-0.14603099243353101");
    REQUIRE(same_bits(got, SERDE_QUIRK_LITERAL_VALUE));
    REQUIRE(!same_bits(got, std::strtod("-0.14603099243353101", nullptr)));
    REQUIRE(same_bits(json::serde_json_f64("0.0"), 0.0) && same_bits(json::serde_json_f64("-0.0"), -0.0));
    REQUIRE(json::serde_json_f64("1.5e3") == 1500.0 && json::serde_json_f64("2E+2") == 200.0);
    REQUIRE(json::serde_json_f64("0.30000000000000004") == 0.30000000000000004);
    REQUIRE(json::serde_json_f64("18446744073709551616") == 1844674407370955161.0 * 10.0);
    REQUIRE(json::serde_json_f64("1e-400") == 0.0);
    for (const char *bad : {"01", "1.", "-", "1e", ".5", "1e400", ""}) {
        bool threw = false;
        try {
            json::serde_json_f64(bad);
        } catch (const json::DeserializationError &) {
            threw = true;
        }
        REQUIRE(threw);
    }
}
TEST(host_deserialize_fixtures_like_the_reference) { // src/simulation.rs:409-420, src/grid/mod.rs:803-823
    for (const ParsedSnap *f : {&FIXTURE_SMALL_SIMULATION, &FIXTURE_SIMPLE_SIMULATION}) {
        std::istringstream reader(f->raw);
        const UnfinalizedSimulation u = json::unfinalized_simulation_from_reader(reader);
        REQUIRE(u.size == (GridSize{f->nx, f->ny}));
        REQUIRE(!u.initial_norm_squared.has_value()); // the files do not carry it
        REQUIRE(u.max_iterations == 100 && u.iterations == 0 && u.omega == 1.7);
        require_parsed(u.grid, *f, "simulation fixture");
    }
    REQUIRE(json::unfinalized_simulation_from_reader(*std::make_unique<std::istringstream>(
                FIXTURE_SMALL_SIMULATION.raw)).cell_size == (CellPhysicalSize{0.1, 0.2}));
    for (const ParsedSnap *f : {&FIXTURE_SMALL_GRID, &FIXTURE_SIMPLE_GRID, &FIXTURE_NAST2D_GRID}) {
        std::istringstream reader(f->raw);
        require_parsed(json::unfinalized_grid_from_reader(reader), *f, "grid fixture");
    }
}
// python/test_generate_test_data.py:28-49: the NaSt2D .out fixture through the converter
TEST(host_nast2d_out_file_like_the_reference_converter) {
    const std::string bytes(reinterpret_cast<const char *>(NAST2D_OUT), sizeof NAST2D_OUT);
    const UnfinalizedSimulationGrid g = nast2d::grid_from_out(bytes);
    require_parsed(g, NAST2D_EXPECTED, "small_data.out");
    // ... and the JSON the reference keeps of the same state (tests/test_data/small_data.out.json)
    // parses to the same grid -- except for the ONE pressure the reference's JSON parser reads
    // an ulp low, p(2, 1) = -0.14603099243353101: the binary holds the exact double
    std::istringstream reader(FIXTURE_NAST2D_GRID.raw);
    const UnfinalizedSimulationGrid j = json::unfinalized_grid_from_reader(reader);
    REQUIRE(j.u == g.u && j.v == g.v && j.cell_type == g.cell_type);
    for (std::size_t x = 0; x < g.size[0]; ++x)
        for (std::size_t y = 0; y < g.size[1]; ++y) {
            if (x == 2 && y == 1) {
                REQUIRE(g.pressure(x, y) == std::strtod("-0.14603099243353101", nullptr));
                REQUIRE(same_bits(j.pressure(x, y), SERDE_QUIRK_LITERAL_VALUE));
            } else {
                REQUIRE(same_bits(j.pressure(x, y), g.pressure(x, y)));
            }
        }
    bool threw = false;
    try {
        nast2d::grid_from_out(bytes.substr(0, bytes.size() - 5));
    } catch (const json::DeserializationError &) {
        threw = true;
    }
    REQUIRE(threw);
}
TEST(host_serialize_round_trip) {
    std::istringstream reader(FIXTURE_SMALL_SIMULATION.raw);
    UnfinalizedSimulation u = json::unfinalized_simulation_from_reader(reader);
    u.initial_norm_squared = 899.9547140394143;
    u.time = 1.0050000000000006;
    const std::string text = json::to_json(u);
    REQUIRE(text.find("\"initial_norm_squared\":899.9547140394143,") != std::string::npos);
    REQUIRE(text.find("{\"Boundary\":{\"Inflow\":{\"velocity\":[1.0,0.0]}}}") != std::string::npos);
    REQUIRE(text.rfind("{\"size\":[4,3],\"cell_size\":[0.1,0.2],\"delt\":0.005,", 0) == 0);
    const UnfinalizedSimulation back = json::simulation_from(json::parse(text));
    REQUIRE(back.grid.pressure == u.grid.pressure && back.grid.u == u.grid.u && back.grid.v == u.grid.v);
    REQUIRE(back.grid.cell_type == u.grid.cell_type);
    REQUIRE(back.initial_norm_squared == u.initial_norm_squared && same_bits(back.time, u.time));
    bool threw = false;
    try {
        json::parse("{\"size\": [4, 3]");
    } catch (const json::DeserializationError &e) {
        threw = std::string(e.what()).find("deserializing") != std::string::npos;
    }
    REQUIRE(threw);
}
// src/simulation.rs:409-420 `deserialize`: from_reader on the two fixture files, on the GPU
TEST(simulation_deserialize) {
    for (const ParsedSnap *f : {&FIXTURE_SIMPLE_SIMULATION, &FIXTURE_SMALL_SIMULATION}) {
        std::istringstream reader(f->raw);
        Simulation sim = simulation_from_reader(reader);
        const double want = f == &FIXTURE_SMALL_SIMULATION ? FIXTURE_SMALL_SIMULATION_INITIAL_NORM
                                                           : FIXTURE_SIMPLE_SIMULATION_INITIAL_NORM;
        REQUIRE(sim.initial_norm_squared().has_value());
        REQUIRE(close(*sim.initial_norm_squared(), want));   // 899.9547140394143 / 0.0
        const UnfinalizedSimulation out = sim.to_unfinalized();
        require_parsed(out.grid, *f, "state after from_reader");
        const UnfinalizedSimulation again = json::simulation_from(json::parse(simulation_to_json(sim)));
        REQUIRE(again.grid.pressure == out.grid.pressure && again.grid.cell_type == out.grid.cell_type);
    }
}

// ---- ticks on HOST arrays, several in flight (include/stroemung_b200_pipeline.hpp) ------------
// five independent states through a pipeline of three handles, four ticks each, on pinned
// buffers: every request ends on the oracle's bits whatever handle served which step
TEST(pipeline_three_requests_in_flight) {
    const GridSize size{330, 420}; // large enough for the pass kernels
    const std::size_t n = size[0] * size[1];
    UnfinalizedSimulation base = tick_case(size, presets::obstacle(size));
    // the residual norm a solve has to beat is latched state of a HANDLE
    // (src/simulation.rs:229-237): every handle and every oracle gets the same one
    base.initial_norm_squared = 2.5e-3;
    base.max_iterations = 40; // ten passes of four sweeps per tick
    Extensions ext;
    ext.sor_mode = SB_SOR_RED_BLACK;
    HostPipeline pipe(3, [&] { return Simulation::try_from(base, ext); });
    REQUIRE(pipe.depth() == 3 && pipe.field_len() == n);

    std::vector<std::uint8_t> kind;
    std::vector<sb_boundary_velocity> tab;
    SimulationGrid::flatten(base.grid.cell_type, kind, tab);
    std::vector<double> bu(n, 0.0), bv(n, 0.0);
    for (const sb_boundary_velocity &t : tab) {
        bu[t.x * size[1] + t.y] = t.u;
        bv[t.x * size[1] + t.y] = t.v;
    }
    const int requests = 5, ticks = 4;
    std::vector<PinnedField> fields; // p, u, v of request r at 3 r, 3 r + 1, 3 r + 2
    std::vector<so_sim *> oracles;
    std::uint64_t rng = 0x5EED5EEDull;
    auto uniform = [&rng] { // splitmix64 -> [-1, 1)
        rng += 0x9E3779B97F4A7C15ull;
        std::uint64_t z = rng;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        return static_cast<double>(z >> 11) * (2.0 / 9007199254740992.0) - 1.0;
    };
    for (int r = 0; r < requests; ++r) {
        for (int k = 0; k < 3; ++k) {
            fields.emplace_back(n);
            for (std::size_t i = 0; i < n; ++i) fields.back()[i] = (k == 0 ? 1.0 : 0.05) * uniform();
        }
        so_params op{};
        op.nx = size[0];
        op.ny = size[1];
        op.delx = base.cell_size[0];
        op.dely = base.cell_size[1];
        op.delt = base.delt;
        op.gamma = base.gamma;
        op.reynolds = base.reynolds;
        op.sor_absolute_epsilon = base.sor_absolute_epsilon;
        op.omega = base.omega;
        op.max_iterations = base.max_iterations;
        op.has_initial_norm = 1;
        op.initial_norm_squared = 2.5e-3;
        op.sor_mode = SO_SOR_RED_BLACK;
        so_sim *o = nullptr;
        std::uint64_t err[2];
        REQUIRE(so_create(&op, fields[3 * r].data(), fields[3 * r + 1].data(), fields[3 * r + 2].data(),
                          kind.data(), bu.data(), bv.data(), &o, err) == SO_OK);
        oracles.push_back(o);
    }
    std::vector<std::vector<HostPipeline::Result>> results(requests);
    for (int t = 0; t < ticks; ++t) {
        std::vector<std::future<HostPipeline::Result>> futs; // 5 requests queued on 3 handles
        for (int r = 0; r < requests; ++r)
            futs.push_back(pipe.submit(fields[3 * r].data(), fields[3 * r + 1].data(), fields[3 * r + 2].data()));
        for (int r = 0; r < requests; ++r) results[r].push_back(futs[r].get());
    }
    for (int r = 0; r < requests; ++r) {
        for (int t = 0; t < ticks; ++t) {
            std::uint32_t oit = 0;
            double onorm = 0.0;
            REQUIRE(so_tick(oracles[r], &oit, &onorm) == SO_OK);
            REQUIRE(results[r][t].first == oit);
            REQUIRE(close(results[r][t].second, onorm));
        }
        for (std::size_t i = 0; i < n; ++i) {
            REQUIRE(same_bits(fields[3 * r][i], so_p(oracles[r])[i]));
            REQUIRE(same_bits(fields[3 * r + 1][i], so_u(oracles[r])[i]));
            REQUIRE(same_bits(fields[3 * r + 2][i], so_v(oracles[r])[i]));
        }
        so_destroy(oracles[r]);
    }
    // a failing request surfaces through its future, the pipeline keeps serving
    bool threw = false;
    try {
        pipe.submit(nullptr, nullptr, nullptr).get();
    } catch (const SimulationError &) {
        threw = true;
    }
    REQUIRE(threw);
    REQUIRE(pipe.submit(fields[0].data(), fields[1].data(), fields[2].data()).get().first > 0);
    pipe.close();
}

// ---- no CPU fallback: a construction that cannot reach a GPU is an error, not a slow answer ----
TEST(invalid_arguments_are_errors) {
    const GridSize size{4, 3};
    UnfinalizedSimulation u = tick_case(size, presets::simple_inflow(size));
    u.grid.pressure = GridArray<Real>::zeros({3, 3});
    bool threw = false;
    try {
        Simulation::try_from(u);
    } catch (const InvalidArgument &) {
        threw = true;
    }
    REQUIRE(threw);
}

int main(int argc, char **argv) {
    std::vector<std::string> only;
    for (int i = 1; i < argc; ++i) {
        if (!std::strcmp(argv[i], "--host")) {
            for (auto &t : registry())
                if (t.first.rfind("host_", 0) == 0) only.push_back(t.first);
            continue;
        }
        if (!std::strcmp(argv[i], "--list")) {
            for (auto &t : registry()) std::printf("%s\n", t.first.c_str());
            std::printf("library: %s\n", sb_version());
            return 0;
        }
        only.emplace_back(argv[i]);
    }
    int failed = 0, ran = 0;
    for (auto &t : registry()) {
        if (!only.empty() && std::find(only.begin(), only.end(), t.first) == only.end()) continue;
        ++ran;
        try {
            t.second();
            std::printf("ok      %s\n", t.first.c_str());
        } catch (const Failure &f) {
            ++failed;
            std::printf("FAILED  %s: %s\n", t.first.c_str(), f.what.c_str());
        } catch (const std::exception &e) {
            ++failed;
            std::printf("FAILED  %s: exception: %s\n", t.first.c_str(), e.what());
        }
        std::fflush(stdout);
    }
    std::printf("%d run, %d failed\n", ran, failed);
    return failed ? 1 : (ran ? 0 : 2);
}
