#!/usr/bin/env python
"""Generate golden fixtures from the UNMODIFIED reference.

Runs only in the build container (needs /root/reference).  The reference's
relagn.py is imported as-is under the stub `xspec`/`astropy` modules of
ref_stubs.py; at the kyconv seam (relagn.py:1001,1098,1356) the stub forwards
to the CPU oracle's kyconv restatement (KYCONV itself is not installable
offline -> the relativistic goldens pin the reference's DRIVER around the seam,
not KYCONV's numbers: "parity unpinned" for the seam itself).

Outputs (committed):
  tests/golden/golden_nonrel.npz  attrs + non-relativistic components (W/Hz)
  tests/golden/golden_rel.npz     relativistic components through reference
                                  driver + oracle kyconv, with the tables used
  tests/golden/golden_nthcomp.npz pyNTHCOMP.donthcomp outputs
  tests/golden/golden_units.npz   getter outputs in every unit mode

Usage:  python tests/golden/make_golden.py
"""
import os
import sys
import io
import contextlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_stubs  # noqa: E402
from oracle import oracle as O  # noqa: E402

REF = "/root/reference/src/python_version"

ATTRS = ["risco", "r_sg", "eta", "L_edd", "Mdot_edd", "Rg", "r_h", "r_w", "r_out", "hmax", "gamma_h", "kTe_h", "kTe_w",
         "gamma_w"]

RELAGN_KEYS = ["M", "dist", "log_mdot", "a", "cos_inc", "kTe_hot", "kTe_warm", "gamma_hot", "gamma_warm", "r_hot",
               "r_warm", "log_rout", "fcol", "h_max", "z"]
RELQSO_KEYS = ["M", "dist", "log_mdot", "a", "cos_inc", "fcol", "z"]
RELAGN_DEF = [1e8, 100, -1, 0, 0.5, 100, 0.2, 1.7, 2.7, 10, 20, -1, 1, 10, 0]
RELQSO_DEF = [1e5, 1, -1, 0, 0.5, 1, 0]


def relagn_cases():
    d = dict(zip(RELAGN_KEYS, RELAGN_DEF))
    cases = [("default", {}), ("a998", dict(a=0.998)), ("fcolneg", dict(fcol=-1, M=1e6, log_mdot=-0.5)),
             ("nohot", dict(r_hot=-1)), ("nowarm", dict(r_warm=10, r_hot=10)), ("rout3", dict(log_rout=3.5, M=1e6)),
             ("z05", dict(z=0.5)), ("retro", dict(a=-0.5, r_hot=15)), ("faceon", dict(cos_inc=0.95, a=0.7)),
             ("rw_gt_rout", dict(r_warm=5000.0)), ("hmax_clamp", dict(h_max=50.0)),
             ("rh_lt_risco", dict(r_hot=3.0)), ("narrow_warm", dict(r_hot=10, r_warm=10.2)),
             ("struct", dict(r_hot=10, r_warm=100, log_rout=3)), ("nodisc", dict(r_warm=1e4, log_rout=2))]
    rng = np.random.Generator(np.random.PCG64(20240229))
    for i in range(6):  # C4-style draws
        M = 10 ** rng.uniform(5, 10)
        a = rng.uniform(0, 0.998)
        p = dict(M=M, log_mdot=rng.uniform(-1.5, 0.3), a=a, cos_inc=rng.uniform(0.09, 0.998),
                 kTe_hot=rng.uniform(10, 300), kTe_warm=rng.uniform(0.1, 1), gamma_hot=rng.uniform(1.3, 3),
                 gamma_warm=rng.uniform(2, 5))
        risco = O.lib().ro_risco(a)
        rh = risco + rng.uniform(0.5, 100)
        p["r_hot"] = rh
        p["r_warm"] = rh * (1 + rng.uniform(0, 4))
        p["h_max"] = min(10.0, rh)
        cases.append(("c4_%d" % i, p))
    out = []
    for name, kw in cases:
        q = dict(d)
        q.update(kw)
        out.append((name, [float(q[k]) for k in RELAGN_KEYS]))
    return out


def relqso_cases():
    d = dict(zip(RELQSO_KEYS, RELQSO_DEF))
    cases = [("default", {}), ("m5_lo", dict(log_mdot=-1.6)), ("m5_hi", dict(log_mdot=0.5)),
             ("m8", dict(M=1e8, dist=100)), ("m10_lo", dict(M=1e10, dist=100, log_mdot=-1.6)),
             ("m10_hi", dict(M=1e10, dist=100, log_mdot=0.5)), ("a9", dict(a=0.9, M=1e7, cos_inc=0.8)),
             ("fcolneg", dict(fcol=-1, M=1e6))]
    out = []
    for name, kw in cases:
        q = dict(d)
        q.update(kw)
        out.append((name, [float(q[k]) for k in RELQSO_KEYS]))
    return out


def build(mod, model, pars, ebins=None):
    cls = mod.relagn if model == 0 else mod.relqso
    keys = RELAGN_KEYS if model == 0 else RELQSO_KEYS
    with contextlib.redirect_stdout(io.StringIO()):
        obj = cls(**dict(zip(keys, pars)))
        if ebins is not None:
            obj.new_ear(np.array(ebins))
    obj.set_units("SI")
    return obj


def components(obj, rel):
    with contextlib.redirect_stdout(io.StringIO()):
        d = obj.get_DiscComponent(rel=rel)
        w = obj.get_WarmComponent(rel=rel)
        h = obj.get_HotComponent(rel=rel)
        t = obj.get_totSED(rel=rel)
    return np.stack([d, w, h, t])


def main():
    # a small table set in format 2 (neighbouring spin nodes must both hold circular orbits at the ISCO of any model
    # between them, hence the close pair below 0.998)
    tab_spins = np.array([0.0, 0.5, 0.8, 0.95, 0.9875, 0.998])
    tab_thetas = np.deg2rad([20.0, 45.0, 60.0, 75.0])
    NR, NPHI, r_in = 64, 64, 1.2369
    print("generating oracle tables ...", flush=True)
    tab = O.Tables(tab_spins, tab_thetas, NR, NPHI, r_in)

    def ky(energies, params, flux):
        return O.kyconv(tab, energies, params, flux)

    xs = ref_stubs.install(ky)
    sys.path.insert(0, REF)
    import relagn as ref  # the unmodified reference
    import pyNTHCOMP as refnth

    # published known answer (examples/relqso_modSpec.ipynb:81-83)
    q = build(ref, 1, RELQSO_DEF)
    assert q.gamma_h == 1.9803689563797076 and q.r_h == 14.261351436104999 and q.r_w == 28.522702872209997, \
        (q.gamma_h, q.r_h, q.r_w)
    print("notebook known-answer reproduced bit-exactly")

    grid1000 = np.geomspace(1e-4, 1e3, 1001)
    grid_def = np.geomspace(1e-4, 1e4, 2000)

    # ---------------- non-relativistic ----------------
    store = {"grid1000": grid1000, "grid_def": grid_def, "attr_names": np.array(ATTRS)}
    for model, cases in ((0, relagn_cases()), (1, relqso_cases())):
        tag = "relagn" if model == 0 else "relqso"
        names, P, A, S1000, Sdef, extra, Fhot = [], [], [], [], [], [], []
        for name, pars in cases:
            o = build(ref, model, pars, grid1000)
            at = [float(getattr(o, k)) for k in ATTRS]
            nb = [len(o.logr_ad_bins), len(o.logr_wc_bins), len(o.logr_hc_bins)]
            S1000.append(components(o, False))
            # the hot corona's spectral shape (relagn.py:1200-1223): only set when a hot region exists (:356-357)
            Fhot.append(np.array(getattr(o, "Fnu_seed_hot", np.zeros(len(grid1000) - 1)), dtype=np.float64))
            with contextlib.redirect_stdout(io.StringIO()):
                o.hotCorona_lumin()
            extra.append([o.Ldiss, o.Lseed] + nb)
            o2 = build(ref, model, pars, None)
            Sdef.append(components(o2, False))
            names.append(name)
            P.append(pars)
            A.append(at)
            print(tag, name, "ok", nb, flush=True)
        store[tag + "_names"] = np.array(names)
        store[tag + "_params"] = np.array(P)
        store[tag + "_attrs"] = np.array(A)
        store[tag + "_extra"] = np.array(extra)
        store[tag + "_spec1000"] = np.array(S1000)
        store[tag + "_fnu_seed_hot1000"] = np.array(Fhot)
        store[tag + "_specdef"] = np.array(Sdef)
    np.savez_compressed(os.path.join(HERE, "golden_nonrel.npz"), **store)

    # ---------------- relativistic (reference driver + oracle kyconv) ----------------
    rel_store = {"grid1000": grid1000, "tab_spins": tab_spins, "tab_thetas": tab_thetas,
                 "tab_NR": NR, "tab_NPHI": NPHI, "tab_r_in": r_in, "tab_data": np.array(tab.data)}
    rel_relagn = [c for c in relagn_cases() if c[0] in ("default", "a998", "nohot", "nowarm", "rout3", "z05", "c4_0",
                                                         "c4_1", "c4_2")]
    rel_relqso = [c for c in relqso_cases() if c[0] in ("default", "m8", "m5_hi", "m10_lo")]
    for model, cases in ((0, rel_relagn), (1, rel_relqso)):
        tag = "relagn" if model == 0 else "relqso"
        names, P, S = [], [], []
        for name, pars in cases:
            o = build(ref, model, pars, grid1000)
            xs.calls.clear()
            S.append(components(o, True))
            names.append(name)
            P.append(pars)
            print(tag, "rel", name, "kyconv calls:", len(xs.calls), flush=True)
            if name == "default" and model == 0:
                rel_store["first_call_params"] = np.array(xs.calls[0][1])
                rel_store["first_call_nE"] = xs.calls[0][0]
        rel_store[tag + "_names"] = np.array(names)
        rel_store[tag + "_params"] = np.array(P)
        rel_store[tag + "_spec1000"] = np.array(S)
    np.savez_compressed(os.path.join(HERE, "golden_rel.npz"), **rel_store)

    # ---------------- pyNTHCOMP ----------------
    eg = grid_def[:-1] + 0.5 * np.diff(grid_def)
    eg1000 = grid1000[:-1] + 0.5 * np.diff(grid1000)
    nth_par = [(2.7, 0.2, 0.01), (1.7, 100.0, 0.05), (2.5, 0.2, 0.003), (3.5, 1.0, 0.1), (1.3, 300.0, 1.5),
               (2.0, 0.05, 1e-3), (4.9, 0.1, 0.02), (1.9, 50.0, 0.2)]
    nth_out_def = np.array([refnth.donthcomp(eg, [g_, kte, kbb, 0, 0]) for g_, kte, kbb in nth_par])
    nth_out_1000 = np.array([refnth.donthcomp(eg1000, [g_, kte, kbb, 0, 0]) for g_, kte, kbb in nth_par])
    np.savez_compressed(os.path.join(HERE, "golden_nthcomp.npz"), par=np.array(nth_par), egrid_def=eg,
                        egrid_1000=eg1000, out_def=nth_out_def, out_1000=nth_out_1000)

    # ---------------- unit modes (host epilogue, relagn.py:487-564) ----------------
    o = build(ref, 0, RELAGN_DEF, grid1000)
    o.z = 0.0
    ustore = {}
    for unit in ["cgs", "cgs_wave", "SI", "SI_wave", "counts"]:
        for fl in (False, True):
            o.set_units(unit)
            o.set_flux() if fl else o.set_lum()
            ustore["%s_%d" % (unit, fl)] = components(o, False)[3]
    o2 = build(ref, 0, relagn_cases()[6][1], grid1000)  # z = 0.5
    o2.set_units("counts")
    o2.set_flux()
    ustore["z05_counts_1"] = components(o2, False)[3]
    ustore["z05_E_obs"] = o2.E_obs
    ustore["Ledd_cgs"] = np.array(1.39e31 * 1e8 * 1e7)
    np.savez_compressed(os.path.join(HERE, "golden_units.npz"), **ustore)
    print("done")


if __name__ == "__main__":
    main()
