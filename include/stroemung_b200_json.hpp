/*
 * stroemung_b200_json.hpp -- the reference's on-disk JSON format for the C++ host mirror.
 *
 * `Simulation::from_reader` / `SimulationGrid::from_reader` (src/simulation.rs:117-120,
 * src/grid/mod.rs:334-341) read what serde derives for `UnfinalizedSimulation[Grid]`
 * (src/simulation.rs:30-44, src/grid/mod.rs:85-92):
 *   - arrays are ndarray-serde documents {"v": 1, "dim": [nx, ny], "data": [...]}, data
 *     row-major with index (x, y), y contiguous;
 *   - a cell is "Fluid", {"Boundary": "NoSlip"}, {"Boundary": "Outflow"} or
 *     {"Boundary": {"Inflow": {"velocity": [u, v]}}} (src/cell.rs:6-23).
 * and `#[derive(Serialize)] Simulation` (src/simulation.rs:49-69) writes the same document.
 *
 * Numbers are read the way the reference reads them: serde_json 1.0.140 WITHOUT its
 * `float_roundtrip` feature (Cargo.toml:18, Cargo.lock:526-527) -- a u64 significand scaled by
 * one multiplication or division by a power of ten, which is up to 1 ulp off a correctly
 * rounded parse (the fixture literal -0.14603099243353101 is read as -0.146030992433531, and
 * `initial_norm_squared` 899.9547140394143 of the reference's `deserialize` test depends on it).
 * `serde_json_f64` restates that published algorithm; the crate itself is not vendored in the
 * reference.  Host-side only: nothing here touches the device (tests/cpp runs it on the CPU).
 */
#ifndef STROEMUNG_B200_JSON_HPP
#define STROEMUNG_B200_JSON_HPP

#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <istream>
#include <iterator>
#include <limits>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "stroemung_b200.hpp"

namespace stroemung {
namespace json {

/* SimulationError::DeserializationError / SimulationGridError::DeserializationError */
class DeserializationError : public SimulationError {
  public:
    explicit DeserializationError(const std::string &what)
        : SimulationError(SB_INVALID_ARGUMENT, "An error occurred while deserializing: `" + what + "`") {}
};

/* ---- serde_json 1.0.140 src/de.rs: parse_integer / parse_decimal / parse_exponent /
 *      f64_from_parts, default features ------------------------------------------------------ */
namespace detail {
inline bool u64_overflows(std::uint64_t significand, std::uint64_t digit) {
    constexpr std::uint64_t max = std::numeric_limits<std::uint64_t>::max();
    return significand >= max / 10 && (significand > max / 10 || digit > max % 10);
}
inline double pow10_table(int k) { /* POW10[k], 0 <= k <= 308: the correctly rounded literal 1e<k> */
    static const std::vector<double> table = [] {
        std::vector<double> t(309);
        for (int i = 0; i <= 308; ++i) {
            const std::string lit = "1e" + std::to_string(i);
            t[i] = std::strtod(lit.c_str(), nullptr);
        }
        return t;
    }();
    return table[k];
}
inline double f64_from_parts(bool positive, std::uint64_t significand, std::int64_t exponent) {
    double f = static_cast<double>(significand); /* round to nearest even, like `as f64` */
    for (;;) {
        const std::int64_t k = exponent < 0 ? -exponent : exponent;
        if (k <= 308) {
            if (exponent >= 0) {
                f = f * pow10_table(static_cast<int>(k));
                if (std::isinf(f)) throw DeserializationError("number out of range");
            } else {
                f = f / pow10_table(static_cast<int>(k));
            }
            break;
        }
        if (f == 0.0) break;
        if (exponent >= 0) throw DeserializationError("number out of range");
        f = f / 1e308;
        exponent += 308;
    }
    return positive ? f : -f;
}
} // namespace detail

/* the f64 serde_json deserialises the JSON number [p, end) to; *used = characters consumed */
inline double serde_json_f64(const char *p, const char *end, std::size_t *used = nullptr) {
    const char *const begin = p;
    auto digit = [&](const char *q) { return q < end && *q >= '0' && *q <= '9'; };
    auto bad = [&]() -> double { throw DeserializationError("invalid number"); };
    bool positive = true;
    if (p < end && *p == '-') {
        positive = false;
        ++p;
    }
    if (!digit(p)) bad();
    std::uint64_t significand = 0;
    std::int64_t exponent = 0;
    if (*p == '0') {
        ++p;
        if (digit(p)) bad(); /* only one leading zero */
    } else {
        while (digit(p)) {
            const std::uint64_t d = static_cast<std::uint64_t>(*p - '0');
            if (detail::u64_overflows(significand, d)) {
                while (digit(p)) { /* parse_long_integer: dropped digits still scale the value */
                    ++exponent;
                    ++p;
                }
                break;
            }
            significand = significand * 10 + d;
            ++p;
        }
    }
    if (p < end && *p == '.') { /* parse_decimal */
        ++p;
        const char *const start = p;
        while (digit(p)) {
            const std::uint64_t d = static_cast<std::uint64_t>(*p - '0');
            if (detail::u64_overflows(significand, d)) {
                while (digit(p)) ++p; /* parse_decimal_overflow: further digits are ignored */
                break;
            }
            significand = significand * 10 + d;
            --exponent;
            ++p;
        }
        if (p == start) bad();
    }
    if (p < end && (*p == 'e' || *p == 'E')) { /* parse_exponent */
        ++p;
        bool positive_exp = true;
        if (p < end && (*p == '+' || *p == '-')) {
            positive_exp = *p == '+';
            ++p;
        }
        if (!digit(p)) bad();
        constexpr std::int64_t i32max = std::numeric_limits<std::int32_t>::max();
        std::int64_t exp = 0;
        bool overflow = false;
        while (digit(p)) {
            const std::int64_t d = *p - '0';
            if (exp >= i32max / 10 && (exp > i32max / 10 || d > i32max % 10)) {
                overflow = true; /* parse_exponent_overflow */
                while (digit(p)) ++p;
                break;
            }
            exp = exp * 10 + d;
            ++p;
        }
        if (overflow) {
            if (significand != 0 && positive_exp) throw DeserializationError("number out of range");
            if (used) *used = static_cast<std::size_t>(p - begin);
            return positive ? 0.0 : -0.0;
        }
        exponent = positive_exp ? std::min<std::int64_t>(exponent + exp, i32max)
                                : std::max<std::int64_t>(exponent - exp, -i32max - 1);
    }
    if (used) *used = static_cast<std::size_t>(p - begin);
    return detail::f64_from_parts(positive, significand, exponent);
}
inline double serde_json_f64(const std::string &text) {
    std::size_t used = 0;
    const double v = serde_json_f64(text.data(), text.data() + text.size(), &used);
    if (used != text.size()) throw DeserializationError("invalid number");
    return v;
}

/* ---- a small JSON document model (what the format above needs) --------------------------- */
struct Value {
    enum class Type { Null, Bool, Number, String, Array, Object } type = Type::Null;
    bool boolean = false;
    double number = 0.0;
    std::string string;
    std::vector<Value> array;
    std::vector<std::pair<std::string, Value>> object; /* insertion order kept */

    const Value *find(const std::string &key) const {
        for (const auto &kv : object)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
    const Value &at(const std::string &key) const {
        if (type != Type::Object) throw DeserializationError("expected an object holding `" + key + "`");
        const Value *v = find(key);
        if (!v) throw DeserializationError("missing field `" + key + "`");
        return *v;
    }
};

class Parser {
  public:
    explicit Parser(const std::string &text) : p_(text.data()), end_(text.data() + text.size()) {}
    Value parse_document() {
        Value v = value();
        ws();
        if (p_ != end_) fail("trailing characters");
        return v;
    }

  private:
    const char *p_, *end_;
    [[noreturn]] void fail(const std::string &what) const { throw DeserializationError(what); }
    void ws() {
        while (p_ < end_ && (*p_ == ' ' || *p_ == '\n' || *p_ == '\t' || *p_ == '\r')) ++p_;
    }
    bool eat(char c) {
        ws();
        if (p_ < end_ && *p_ == c) {
            ++p_;
            return true;
        }
        return false;
    }
    void literal(const char *word) {
        for (const char *w = word; *w; ++w, ++p_)
            if (p_ >= end_ || *p_ != *w) fail("invalid literal");
    }
    std::string string() {
        std::string out;
        ++p_; /* opening quote */
        while (p_ < end_ && *p_ != '"') {
            if (*p_ == '\\') {
                if (++p_ >= end_) fail("EOF while parsing a string");
                switch (*p_) {
                case 'n': out += '\n'; break;
                case 't': out += '\t'; break;
                case 'r': out += '\r'; break;
                case 'b': out += '\b'; break;
                case 'f': out += '\f'; break;
                case 'u': fail("\\u escapes do not occur in this format");
                default: out += *p_;
                }
                ++p_;
            } else {
                out += *p_++;
            }
        }
        if (p_ >= end_) fail("EOF while parsing a string");
        ++p_;
        return out;
    }
    Value value() {
        ws();
        if (p_ >= end_) fail("EOF while parsing a value");
        Value v;
        const char c = *p_;
        if (c == '{') {
            ++p_;
            v.type = Value::Type::Object;
            if (eat('}')) return v;
            do {
                ws();
                if (p_ >= end_ || *p_ != '"') fail("key must be a string");
                std::string key = string();
                if (!eat(':')) fail("expected `:`");
                v.object.emplace_back(std::move(key), value());
            } while (eat(','));
            if (!eat('}')) fail("expected `,` or `}`");
        } else if (c == '[') {
            ++p_;
            v.type = Value::Type::Array;
            if (eat(']')) return v;
            do v.array.push_back(value());
            while (eat(','));
            if (!eat(']')) fail("expected `,` or `]`");
        } else if (c == '"') {
            v.type = Value::Type::String;
            v.string = string();
        } else if (c == 't') {
            literal("true");
            v.type = Value::Type::Bool;
            v.boolean = true;
        } else if (c == 'f') {
            literal("false");
            v.type = Value::Type::Bool;
        } else if (c == 'n') {
            literal("null");
        } else {
            std::size_t used = 0;
            v.type = Value::Type::Number;
            v.number = serde_json_f64(p_, end_, &used);
            p_ += used;
        }
        return v;
    }
};

inline Value parse(const std::string &text) { return Parser(text).parse_document(); }

/* ---- the reference's structs ---------------------------------------------------------------- */
namespace detail {
inline double number(const Value &v, const char *what) {
    if (v.type != Value::Type::Number) throw DeserializationError(std::string("expected a number for ") + what);
    return v.number;
}
inline std::uint64_t integer(const Value &v, const char *what) {
    const double d = number(v, what);
    if (d < 0 || d != std::floor(d)) throw DeserializationError(std::string("expected an unsigned integer for ") + what);
    return static_cast<std::uint64_t>(d);
}
template <std::size_t N>
inline std::array<double, N> reals(const Value &v, const char *what) {
    if (v.type != Value::Type::Array || v.array.size() != N)
        throw DeserializationError(std::string("expected an array of ") + std::to_string(N) + " for " + what);
    std::array<double, N> out{};
    for (std::size_t i = 0; i < N; ++i) out[i] = number(v.array[i], what);
    return out;
}
inline GridSize size_of(const Value &v, const char *what) {
    if (v.type != Value::Type::Array || v.array.size() != 2)
        throw DeserializationError(std::string("expected [nx, ny] for ") + what);
    return {static_cast<std::size_t>(integer(v.array[0], what)), static_cast<std::size_t>(integer(v.array[1], what))};
}
/* ndarray-serde: {"v": 1, "dim": [nx, ny], "data": [...]} */
inline const Value &array_body(const Value &doc, GridSize &dim, const char *what) {
    if (integer(doc.at("v"), "v") != 1) throw DeserializationError("unknown array version");
    dim = size_of(doc.at("dim"), "dim");
    const Value &data = doc.at("data");
    if (data.type != Value::Type::Array || data.array.size() != dim[0] * dim[1])
        throw DeserializationError(std::string("data length does not match dim for ") + what);
    return data;
}
inline Cell cell_of(const Value &v) { /* src/cell.rs:6-23 as serde's externally tagged enums */
    if (v.type == Value::Type::String) {
        if (v.string == "Fluid") return Cell::Fluid();
        throw DeserializationError("unknown variant `" + v.string + "`, expected `Fluid` or `Boundary`");
    }
    const Value &b = v.at("Boundary");
    if (b.type == Value::Type::String) {
        if (b.string == "NoSlip") return Cell::Boundary(BoundaryCell::NoSlip());
        if (b.string == "Outflow") return Cell::Boundary(BoundaryCell::Outflow());
        throw DeserializationError("unknown variant `" + b.string + "`");
    }
    if (const Value *in = b.find("Inflow"))
        return Cell::Boundary(BoundaryCell::Inflow(reals<2>(in->at("velocity"), "velocity")));
    if (const Value *mw = b.find("MovingWall")) /* extension kind of this build */
        return Cell::Boundary(BoundaryCell::MovingWall(reals<2>(mw->at("velocity"), "velocity")));
    throw DeserializationError("unknown BoundaryCell variant");
}
} // namespace detail

inline GridArray<Real> real_array_from(const Value &doc, const char *what) {
    GridSize dim{};
    const Value &data = detail::array_body(doc, dim, what);
    GridArray<Real> a(dim);
    for (std::size_t i = 0; i < data.array.size(); ++i) a.data()[i] = detail::number(data.array[i], what);
    return a;
}
inline GridArray<Cell> cell_array_from(const Value &doc) {
    GridSize dim{};
    const Value &data = detail::array_body(doc, dim, "cell_type");
    GridArray<Cell> a(dim);
    for (std::size_t i = 0; i < data.array.size(); ++i) a.data()[i] = detail::cell_of(data.array[i]);
    return a;
}

/* UnfinalizedSimulationGrid (src/grid/mod.rs:85-92) */
inline UnfinalizedSimulationGrid grid_from(const Value &doc) {
    UnfinalizedSimulationGrid g;
    g.size = detail::size_of(doc.at("size"), "size");
    g.pressure = real_array_from(doc.at("pressure"), "pressure");
    g.u = real_array_from(doc.at("u"), "u");
    g.v = real_array_from(doc.at("v"), "v");
    g.cell_type = cell_array_from(doc.at("cell_type"));
    return g;
}
/* UnfinalizedSimulation (src/simulation.rs:30-44); `initial_norm_squared: Option<Real>` may be
 * absent or null */
inline UnfinalizedSimulation simulation_from(const Value &doc) {
    UnfinalizedSimulation u;
    u.size = detail::size_of(doc.at("size"), "size");
    u.cell_size = detail::reals<2>(doc.at("cell_size"), "cell_size");
    u.delt = detail::number(doc.at("delt"), "delt");
    u.gamma = detail::number(doc.at("gamma"), "gamma");
    u.reynolds = detail::number(doc.at("reynolds"), "reynolds");
    if (const Value *n = doc.find("initial_norm_squared"))
        if (n->type != Value::Type::Null) u.initial_norm_squared = detail::number(*n, "initial_norm_squared");
    u.sor_absolute_epsilon = detail::number(doc.at("sor_absolute_epsilon"), "sor_absolute_epsilon");
    u.max_iterations = static_cast<std::uint32_t>(detail::integer(doc.at("max_iterations"), "max_iterations"));
    u.iterations = static_cast<std::uint32_t>(detail::integer(doc.at("iterations"), "iterations"));
    u.time = detail::number(doc.at("time"), "time");
    u.omega = detail::number(doc.at("omega"), "omega");
    u.grid = grid_from(doc.at("grid"));
    return u;
}

inline std::string slurp(std::istream &reader) {
    return std::string(std::istreambuf_iterator<char>(reader), std::istreambuf_iterator<char>());
}
inline UnfinalizedSimulation unfinalized_simulation_from_reader(std::istream &reader) {
    return simulation_from(parse(slurp(reader)));
}
inline UnfinalizedSimulationGrid unfinalized_grid_from_reader(std::istream &reader) {
    return grid_from(parse(slurp(reader)));
}

/* ---- Serialize: what `serde_json::to_string(&simulation)` writes --------------------------- */
inline std::string real_to_json(Real x) { /* shortest literal that round-trips, like ryu */
    if (!std::isfinite(x)) return "null"; /* serde_json writes null for NaN / inf */
    char buf[40];
    auto r = std::to_chars(buf, buf + sizeof buf, x);
    std::string s(buf, r.ptr);
    if (s.find_first_of(".eEn") == std::string::npos) s += ".0"; /* 1 -> 1.0 like serde_json */
    return s;
}
inline std::string to_json(const GridArray<Real> &a) {
    std::string s = "{\"v\":1,\"dim\":[" + std::to_string(a.dim()[0]) + "," + std::to_string(a.dim()[1]) + "],\"data\":[";
    for (std::size_t i = 0; i < a.len(); ++i) s += (i ? "," : "") + real_to_json(a.data()[i]);
    return s + "]}";
}
inline std::string to_json(const Cell &c) {
    if (c.is_fluid()) return "\"Fluid\"";
    const BoundaryCell &b = c.boundary();
    switch (b.kind) {
    case BoundaryCell::Kind::NoSlip: return "{\"Boundary\":\"NoSlip\"}";
    case BoundaryCell::Kind::Outflow: return "{\"Boundary\":\"Outflow\"}";
    case BoundaryCell::Kind::Inflow:
        return "{\"Boundary\":{\"Inflow\":{\"velocity\":[" + real_to_json(b.velocity[0]) + "," + real_to_json(b.velocity[1]) + "]}}}";
    default:
        return "{\"Boundary\":{\"MovingWall\":{\"velocity\":[" + real_to_json(b.velocity[0]) + "," + real_to_json(b.velocity[1]) + "]}}}";
    }
}
inline std::string to_json(const GridArray<Cell> &a) {
    std::string s = "{\"v\":1,\"dim\":[" + std::to_string(a.dim()[0]) + "," + std::to_string(a.dim()[1]) + "],\"data\":[";
    for (std::size_t i = 0; i < a.len(); ++i) s += (i ? "," : "") + to_json(a.data()[i]);
    return s + "]}";
}
inline std::string to_json(const UnfinalizedSimulationGrid &g) { /* field order of src/grid/mod.rs:112-117 */
    return "{\"size\":[" + std::to_string(g.size[0]) + "," + std::to_string(g.size[1]) + "],\"pressure\":" +
           to_json(g.pressure) + ",\"u\":" + to_json(g.u) + ",\"v\":" + to_json(g.v) + ",\"cell_type\":" +
           to_json(g.cell_type) + "}";
}
inline std::string to_json(const UnfinalizedSimulation &u) { /* field order of src/simulation.rs:49-69 */
    return "{\"size\":[" + std::to_string(u.size[0]) + "," + std::to_string(u.size[1]) + "],\"cell_size\":[" +
           real_to_json(u.cell_size[0]) + "," + real_to_json(u.cell_size[1]) + "],\"delt\":" + real_to_json(u.delt) +
           ",\"gamma\":" + real_to_json(u.gamma) + ",\"reynolds\":" + real_to_json(u.reynolds) +
           ",\"initial_norm_squared\":" + (u.initial_norm_squared ? real_to_json(*u.initial_norm_squared) : "null") +
           ",\"sor_absolute_epsilon\":" + real_to_json(u.sor_absolute_epsilon) +
           ",\"max_iterations\":" + std::to_string(u.max_iterations) + ",\"iterations\":" + std::to_string(u.iterations) +
           ",\"time\":" + real_to_json(u.time) + ",\"omega\":" + real_to_json(u.omega) + ",\"grid\":" + to_json(u.grid) + "}";
}

} // namespace json

/* ---- NaSt2D `.out` binaries (what the reference's converter reads,
 * python/generate_test_data.py:107-162): two C ints imax, jmax; then U, V, P, T as
 * (imax+2)*(jmax+2) C doubles each, x-major; then the flag field as C ints (bit 0x10 = fluid).
 * The converter's boundary-kind reconstruction (:46-78): a non-fluid cell of the left wall is
 * Inflow carrying the file's (u, v), of the right wall Outflow, everything else NoSlip. ------- */
namespace nast2d {
inline UnfinalizedSimulationGrid grid_from_out(const std::string &bytes) {
    auto need = [&](std::size_t n) {
        if (bytes.size() < n) throw json::DeserializationError("NaSt2D .out file is truncated");
    };
    auto int_at = [&](std::size_t off) {
        std::int32_t v;
        std::memcpy(&v, bytes.data() + off, 4); /* little-endian C int, like the converter's default */
        return v;
    };
    need(8);
    const std::int32_t imax = int_at(0), jmax = int_at(4);
    if (imax < 0 || jmax < 0) throw json::DeserializationError("NaSt2D .out file: negative size");
    const GridSize size{static_cast<std::size_t>(imax) + 2, static_cast<std::size_t>(jmax) + 2};
    const std::size_t n = size[0] * size[1];
    need(8 + 4 * n * 8 + n * 4);
    UnfinalizedSimulationGrid g;
    g.size = size;
    GridArray<Real> *fields[3] = {&g.u, &g.v, &g.pressure}; /* file order U, V, P, (T) */
    for (int f = 0; f < 3; ++f) {
        *fields[f] = GridArray<Real>(size);
        std::memcpy(fields[f]->data(), bytes.data() + 8 + static_cast<std::size_t>(f) * n * 8, n * 8);
    }
    const std::size_t flags = 8 + 4 * n * 8;
    g.cell_type = GridArray<Cell>(size, Cell::Boundary(BoundaryCell::NoSlip()));
    auto fluid = [&](std::size_t x, std::size_t y) { return (int_at(flags + (x * size[1] + y) * 4) & 0x10) != 0; };
    for (std::size_t x = 0; x < size[0]; ++x)
        for (std::size_t y = 0; y < size[1]; ++y)
            if (fluid(x, y)) g.cell_type(x, y) = Cell::Fluid();
    for (std::size_t y = 1; y + 1 < size[1]; ++y) {
        if (!fluid(0, y)) g.cell_type(0, y) = Cell::Boundary(BoundaryCell::Inflow({g.u(0, y), g.v(0, y)}));
        if (!fluid(size[0] - 1, y)) g.cell_type(size[0] - 1, y) = Cell::Boundary(BoundaryCell::Outflow());
    }
    return g;
}
inline UnfinalizedSimulationGrid grid_from_reader(std::istream &reader) { return grid_from_out(json::slurp(reader)); }
} // namespace nast2d

/* Simulation::from_reader (src/simulation.rs:117-120): deserialise, then try_from */
inline Simulation simulation_from_reader(std::istream &reader, const Extensions &ext = {}) {
    return Simulation::try_from(json::unfinalized_simulation_from_reader(reader), ext);
}
/* `serde_json::to_string(&simulation)`: the device state as the reference's document */
inline std::string simulation_to_json(const Simulation &sim) { return json::to_json(sim.to_unfinalized()); }

} // namespace stroemung

#endif /* STROEMUNG_B200_JSON_HPP */
