"""A second, independent restatement of the reference's tick -- pure Python floats and loops.

Test infrastructure only (like oracle/): written straight from the Rust sources, without
looking at oracle/stroemung_oracle.c, so that the C oracle can be cross-checked on cases the
reference's own tests never reach (grids beyond 5x7, obstacles, every corner edge class,
interior Inflow / Outflow cells; SURVEY.md section 8c "coverage gaps").  Python floats are
IEEE doubles, `a*b+c` is never contracted, `/` is the IEEE division and `x.powi(2)` is `x*x`:
the arithmetic below rounds exactly like rustc's output for the expressions it copies.

Index convention of the reference: arrays are [x][y], (0, 0) is the upper-left corner, north
is y-1 (src/grid/mod.rs:167-200).
"""
import math

FLUID, NOSLIP, OUTFLOW, INFLOW = 0, 1, 2, 3
# edge classes, numbered like include/stroemung_b200.h
NONE, N, NE, E, SE, S, SW, W, NW = range(9)


class TooThin(Exception):
    pass


def classify(kind, nx, ny):
    """rebuild_boundary_list + calculate_edges (src/grid/mod.rs:202-235, 270-332):
    -> (sorted list of ((x, y), edge), fluid cell count)"""
    def fluid(x, y):
        return 0 <= x < nx and 0 <= y < ny and kind[x][y] == FLUID
    table = {  # (left, right, up, down) -> edge, src/grid/mod.rs:297-331
        (0, 0, 0, 0): NONE, (1, 0, 0, 0): W, (1, 0, 1, 0): NW, (0, 0, 1, 0): N, (0, 1, 1, 0): NE,
        (0, 1, 0, 0): E, (0, 1, 0, 1): SE, (0, 0, 0, 1): S, (1, 0, 0, 1): SW,
    }
    lst, fluid_cells = [], 0
    for x in range(nx):            # BTreeSet<BoundaryIndex> order: x-major (src/types.rs:12-19)
        for y in range(ny):
            if kind[x][y] == FLUID:
                fluid_cells += 1
                continue
            key = (int(fluid(x - 1, y)), int(fluid(x + 1, y)), int(fluid(x, y - 1)),
                   int(fluid(x, y + 1)))
            if key not in table:
                raise TooThin((x, y))
            lst.append(((x, y), table[key]))
    return lst, float(fluid_cells)


def du2dx(u, i, j, delx, gamma):          # src/math.rs:19-33
    um, uc, up = u[i - 1][j], u[i][j], u[i + 1][j]
    left = ((uc + up) * (uc + up)) - ((um + uc) * (um + uc))
    return (left + (gamma * ((abs(uc + up) * (uc - up)) - (abs(um + uc) * (um - uc))))) / (4.0 * delx)


def duvdx(u, v, i, j, delx, gamma):       # src/math.rs:53-77
    u_ij, u_ijp, u_imj, u_imjp = u[i][j], u[i][j + 1], u[i - 1][j], u[i - 1][j + 1]
    v_ij, v_ipj, v_imj = v[i][j], v[i + 1][j], v[i - 1][j]
    left = ((u_ij + u_ijp) * (v_ij + v_ipj)) - ((u_imj + u_imjp) * (v_imj + v_ij))
    il2 = abs(u_ij + u_ijp) * (v_ij - v_ipj)
    ir2 = abs(u_imj + u_imjp) * (v_imj - v_ij)
    return (left + (gamma * (il2 - ir2))) / (4.0 * delx)


def duvdy(u, v, i, j, dely, gamma):       # src/math.rs:97-120
    u_ij, u_ijm, u_ijp = u[i][j], u[i][j - 1], u[i][j + 1]
    v_ij, v_ijm, v_ipj, v_ipjm = v[i][j], v[i][j - 1], v[i + 1][j], v[i + 1][j - 1]
    left = ((v_ij + v_ipj) * (u_ij + u_ijp)) - ((v_ijm + v_ipjm) * (u_ijm + u_ij))
    il2 = abs(v_ij + v_ipj) * (u_ij - u_ijp)
    ir2 = abs(v_ijm + v_ipjm) * (u_ijm - u_ij)
    return (left + (gamma * (il2 - ir2))) / (4.0 * dely)


def dv2dy(v, i, j, dely, gamma):          # src/math.rs:136-150
    vc, vp, vm = v[i][j], v[i][j + 1], v[i][j - 1]
    left = ((vc + vp) * (vc + vp)) - ((vm + vc) * (vm + vc))
    return (left + (gamma * ((abs(vc + vp) * (vc - vp)) - (abs(vm + vc) * (vm - vc))))) / (4.0 * dely)


def laplacian(e, i, j, delx, dely):       # src/math.rs:162-174
    d2x = ((e[i + 1][j] - (2.0 * e[i][j])) + e[i - 1][j]) / (delx * delx)
    d2y = ((e[i][j + 1] - (2.0 * e[i][j])) + e[i][j - 1]) / (dely * dely)
    return d2x + d2y


def residual(p, i, j, delx, dely, rhs):   # src/math.rs:176-186
    part1 = ((p[i + 1][j] - p[i][j]) - (p[i][j] - p[i - 1][j])) / (delx * delx)
    part2 = ((p[i][j + 1] - p[i][j]) - (p[i][j] - p[i][j - 1])) / (dely * dely)
    return (part1 + part2) - rhs


class PySim:
    """Simulation (src/simulation.rs:49-69) with plain lists of lists."""

    def __init__(self, nx, ny, kind, bu, bv, p, u, v, *, delx, dely, delt, gamma, reynolds,
                 sor_absolute_epsilon, max_iterations, omega, initial_norm_squared=None):
        self.nx, self.ny = nx, ny
        cp = lambda a: [[float(a[x][y]) for y in range(ny)] for x in range(nx)]
        self.kind = [[int(kind[x][y]) for y in range(ny)] for x in range(nx)]
        self.bu, self.bv, self.p, self.u, self.v = cp(bu), cp(bv), cp(p), cp(u), cp(v)
        zeros = lambda: [[0.0] * ny for _ in range(nx)]
        self.f, self.g, self.rhs = zeros(), zeros(), zeros()
        self.delx, self.dely, self.delt, self.gamma, self.reynolds = delx, dely, delt, gamma, reynolds
        self.eps, self.max_iterations, self.omega = sor_absolute_epsilon, max_iterations, omega
        self.time, self.iterations = 0.0, 0
        self.restore = []
        # try_from (src/simulation.rs:71-99): classify, ranges, F/G, RHS, initial norm -- no
        # velocity BC at construction
        self.blist, self.fluid_cells = classify(self.kind, nx, ny)
        self.calculate_pressure_range()
        self.calculate_speed_range()
        self.calculate_f_and_g()
        self.calculate_rhs()
        self.initial_norm_squared = (self.calculate_norm_squared() if initial_norm_squared is None
                                     else initial_norm_squared)

    # ---- src/grid/mod.rs:237-268 ----
    def calculate_pressure_range(self):
        lo, hi = 1.7976931348623157e308, 0.0
        for x in range(self.nx):
            for y in range(self.ny):
                if self.kind[x][y] == FLUID:
                    lo, hi = min(lo, self.p[x][y]), max(hi, self.p[x][y])
        self.pressure_range = [lo, hi]

    def calculate_speed_range(self):
        lo, hi = 1.7976931348623157e308, 0.0
        for x in range(self.nx):
            for y in range(self.ny):
                if self.kind[x][y] == FLUID:
                    s2 = (self.u[x][y] * self.u[x][y]) + (self.v[x][y] * self.v[x][y])
                    lo, hi = min(lo, s2), max(hi, s2)
        self.speed_range = [math.sqrt(lo), math.sqrt(hi)]

    # ---- src/grid/mod.rs:414-651: sequential, in place, in list order ----
    def set_boundary_u_and_v(self):
        u, v = self.u, self.v
        self.restore = []
        for (x, y), edge in self.blist:
            if edge == NONE:
                self.restore.append(((x, y), u[x][y], v[x][y]))
                continue
            n, s, e, w = (x, y - 1), (x, y + 1), (x + 1, y), (x - 1, y)
            k = self.kind[x][y]
            if k in (NOSLIP, INFLOW):
                b_u, b_v = (0.0, 0.0) if k == NOSLIP else (self.bu[x][y], self.bv[x][y])
                if edge == N:
                    u[x][y] = -u[n[0]][n[1]]
                    v[n[0]][n[1]] = b_v
                elif edge == NE:
                    u[x][y] = b_u
                    v[n[0]][n[1]] = b_v
                    v[x][y] = -v[e[0]][e[1]]
                elif edge == E:
                    u[x][y] = b_u
                    v[x][y] = -v[e[0]][e[1]]
                elif edge == SE:
                    u[x][y] = b_u
                    v[x][y] = b_v
                elif edge == S:
                    u[x][y] = -u[s[0]][s[1]]
                    v[x][y] = b_v
                elif edge == SW:
                    u[w[0]][w[1]] = b_u
                    u[x][y] = -u[s[0]][s[1]]
                    v[x][y] = b_v
                elif edge == W:
                    u[w[0]][w[1]] = b_u
                    v[x][y] = -v[w[0]][w[1]]
                elif edge == NW:
                    u[w[0]][w[1]] = b_u
                    u[x][y] = -u[n[0]][n[1]]
                    v[n[0]][n[1]] = b_v
                    v[x][y] = -v[w[0]][w[1]]
            elif k == OUTFLOW:
                src_u = {N: n, NE: n, E: e, SE: e, S: s, SW: w, W: w, NW: n}[edge]
                src_v = {N: n, NE: e, E: e, SE: s, S: s, SW: s, W: w, NW: w}[edge]
                u[x][y] = u[src_u[0]][src_u[1]]
                v[x][y] = v[src_v[0]][src_v[1]]
            else:
                raise RuntimeError("BoundaryListIncorrectError")
            self.restore.append(((x, y), u[x][y], v[x][y]))
            # the second record is keyed by the BOUNDARY cell's index (:602-648)
            if edge in (N, NE):
                self.restore.append(((x, y), None, v[n[0]][n[1]]))
            elif edge in (SW, W):
                self.restore.append(((x, y), u[w[0]][w[1]], None))
            elif edge == NW:
                self.restore.append(((x, y), u[w[0]][w[1]], v[n[0]][n[1]]))

    # ---- src/simulation.rs:122-202, 349-392 ----
    def calculate_f_and_g(self):
        u, v = self.u, self.v
        for i in range(1, self.nx - 1):
            for j in range(1, self.ny - 1):
                self.f[i][j] = u[i][j] + (self.delt * (
                    ((laplacian(u, i, j, self.delx, self.dely) / self.reynolds)
                     - du2dx(u, i, j, self.delx, self.gamma))
                    - duvdy(u, v, i, j, self.dely, self.gamma)))
                self.g[i][j] = v[i][j] + (self.delt * (
                    ((laplacian(v, i, j, self.delx, self.dely) / self.reynolds)
                     - duvdx(u, v, i, j, self.delx, self.gamma))
                    - dv2dy(v, i, j, self.dely, self.gamma)))
        for (x, y), edge in self.blist:
            self.f[x][y] = u[x][y]
            self.g[x][y] = v[x][y]
            if edge in (N, NW, NE):
                self.g[x][y - 1] = v[x][y - 1]
            if edge in (NW, W, SW):
                self.f[x - 1][y] = u[x - 1][y]

    # ---- src/simulation.rs:204-214 ----
    def calculate_rhs(self):
        for i in range(1, self.nx):
            for j in range(1, self.ny):
                self.rhs[i][j] = (((self.f[i][j] - self.f[i - 1][j]) / self.delx)
                                  + ((self.g[i][j] - self.g[i][j - 1]) / self.dely)) / self.delt

    # ---- src/simulation.rs:216-227 ----
    def calculate_norm_squared(self):
        acc = 0.0
        for i in range(1, self.nx - 1):
            for j in range(1, self.ny - 1):
                r = residual(self.p, i, j, self.delx, self.dely, self.rhs[i][j])
                acc = acc + (r * r)
        return acc / self.fluid_cells

    # ---- src/grid/mod.rs:343-412 ----
    def copy_pressure_to_boundaries(self):
        p = self.p
        for (x, y), edge in self.blist:
            if edge == NONE:
                continue
            pn = lambda: p[x][y - 1]
            ps = lambda: p[x][y + 1]
            pe = lambda: p[x + 1][y]
            pw = lambda: p[x - 1][y]
            if edge == N: p[x][y] = pn()
            elif edge == NE: p[x][y] = (pn() + pe()) / 2.0
            elif edge == E: p[x][y] = pe()
            elif edge == SE: p[x][y] = (ps() + pe()) / 2.0
            elif edge == S: p[x][y] = ps()
            elif edge == SW: p[x][y] = (ps() + pw()) / 2.0
            elif edge == W: p[x][y] = pw()
            elif edge == NW: p[x][y] = (pn() + pw()) / 2.0

    # ---- src/simulation.rs:239-285 ----
    def solve_sor(self):
        delx2, dely2 = self.delx * self.delx, self.dely * self.dely
        omw = 1.0 - self.omega
        middle = self.omega / ((2.0 / delx2) + (2.0 / dely2))
        eps2 = self.eps * self.eps
        p, norm = self.p, 0.0
        for it in range(self.max_iterations):
            self.copy_pressure_to_boundaries()
            for x in range(1, self.nx - 1):
                for y in range(1, self.ny - 1):
                    if self.kind[x][y] == FLUID:
                        p[x][y] = (omw * p[x][y]) + middle * (
                            (((p[x + 1][y] + p[x - 1][y]) / delx2)
                             + ((p[x][y + 1] + p[x][y - 1]) / dely2)) - self.rhs[x][y])
            norm = self.calculate_norm_squared()
            if norm < self.initial_norm_squared or norm < eps2:
                return it + 1, norm
        self.calculate_pressure_range()
        return self.max_iterations, norm

    # ---- src/simulation.rs:287-322 ----
    def set_u_and_v(self):
        p = self.p
        for i in range(self.nx - 1):
            for j in range(self.ny - 1):
                self.u[i][j] = self.f[i][j] - (self.delt / self.delx) * (p[i + 1][j] - p[i][j])
                self.v[i][j] = self.g[i][j] - (self.delt / self.dely) * (p[i][j + 1] - p[i][j])
        for (x, y), ru, rv in self.restore:
            if ru is not None:
                self.u[x][y] = ru
            if rv is not None:
                self.v[x][y] = rv
        self.calculate_speed_range()

    # ---- src/simulation.rs:324-333 ----
    def run_simulation_tick(self):
        self.set_boundary_u_and_v()
        self.calculate_f_and_g()
        self.calculate_rhs()
        it, norm = self.solve_sor()
        self.set_u_and_v()
        self.time += self.delt
        self.iterations += 1
        return it, norm
