"""The oracle's results must not depend on how it is compiled: the reference's f64 semantics
(no contraction, IEEE division, the RB extension's explicit fma()) are in the SOURCE, not in an
optimisation level.  The same driver is built from oracle/stroemung_oracle.c at -O0, -O2, -O3 and
-O3 -march=native (hardware FMA for the explicit fma() calls, libm's exact fma otherwise) and
must print identical bits."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent

DRIVER = r'''
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "stroemung_oracle.h"
static unsigned long long fnv(unsigned long long h, const void *p, size_t n) {
    const unsigned char *b = (const unsigned char *)p;
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}
int main(void) {
    const uint64_t nx = 60, ny = 24, n = nx * ny;
    for (int mode = 0; mode < 2; mode++) {
        uint8_t *kind = calloc(n, 1);
        double *bu = calloc(n, 8), *bv = calloc(n, 8);
        so_preset_obstacle(nx, ny, kind, bu, bv);
        so_params p;
        memset(&p, 0, sizeof p);
        p.nx = nx; p.ny = ny; p.delx = 0.1; p.dely = 0.2; p.delt = 0.005; p.gamma = 0.9;
        p.reynolds = 100.0; p.sor_absolute_epsilon = 1e-3; p.omega = 1.7; p.max_iterations = 60;
        p.sor_mode = mode; p.tau = mode ? 0.5 : 0.0;
        so_sim *s = NULL;
        uint64_t err[2];
        if (so_create(&p, NULL, NULL, NULL, kind, bu, bv, &s, err) != SO_OK) return 1;
        for (int t = 0; t < 4; t++) {
            uint32_t it = 0;
            double norm = 0.0;
            if (so_tick(s, &it, &norm) != SO_OK) return 2;
            unsigned long long h = 14695981039346656037ull;
            h = fnv(h, so_p(s), n * 8); h = fnv(h, so_u(s), n * 8); h = fnv(h, so_v(s), n * 8);
            h = fnv(h, so_f(s), n * 8); h = fnv(h, so_g(s), n * 8); h = fnv(h, so_rhs(s), n * 8);
            so_state st;
            so_get_state(s, &st);
            printf("mode %d tick %d it %u norm %a delt %a range %a %a hash %016llx\n", mode, t, it,
                   norm, st.delt, st.pressure_range[1], st.speed_range[1], h);
        }
        so_destroy(s);
        free(kind); free(bu); free(bv);
    }
    return 0;
}
'''

VARIANTS = [["-O0"], ["-O2", "-fno-inline"], ["-O3"], ["-O3", "-march=native"]]


@pytest.fixture(scope="module")
def outputs(tmp_path_factory):
    d = tmp_path_factory.mktemp("oracle_flags")
    (d / "driver.c").write_text(DRIVER)
    outs = []
    for k, flags in enumerate(VARIANTS):
        exe = d / f"driver{k}"
        r = subprocess.run(["gcc", "-std=c11", "-ffp-contract=off", "-fno-fast-math", *flags,
                            f"-I{ROOT / 'oracle'}", str(d / "driver.c"),
                            str(ROOT / "oracle" / "stroemung_oracle.c"), "-o", str(exe), "-lm"],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, (flags, r.stdout, r.stderr)
        outs.append(r.stdout)
    return outs


def test_oracle_bits_do_not_depend_on_compiler_flags(outputs):
    assert len(outputs[0].splitlines()) == 8
    for flags, out in zip(VARIANTS[1:], outputs[1:]):
        assert out == outputs[0], flags


def test_the_shared_object_the_tests_use_agrees_with_them(outputs):
    """...and so does oracle/_build/libstroemung_oracle.so (the Makefile's -O3 build)"""
    import numpy as np
    from oracle import pyoracle as po
    from stroemung_b200 import presets
    g = presets.obstacle((60, 24))
    o = po.OracleSim(60, 24, delx=0.1, dely=0.2, delt=0.005, gamma=0.9, reynolds=100.0,
                     sor_absolute_epsilon=1e-3, max_iterations=60, omega=1.7, kind=g["kind"],
                     bu=g["bu"], bv=g["bv"], sor_mode=po.SOR_REFERENCE_ORDER)
    lines = [l for l in outputs[0].splitlines() if l.startswith("mode 0")]
    for t in range(4):
        it, norm = o.run_simulation_tick()
        tok = lines[t].split()
        assert int(tok[tok.index("it") + 1]) == it
        assert float.fromhex(tok[tok.index("norm") + 1]) == norm, lines[t]
    assert np.isfinite(o.p).all()
