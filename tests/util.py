import numpy as np

from misaki_render_b200 import capi


def relmse(img, ref):
    """mean over pixels and RGB of (I - R)^2 / (R^2 + 1e-2)."""
    img, ref = np.asarray(img, np.float64)[..., :3], np.asarray(ref, np.float64)[..., :3]
    return float(np.mean((img - ref) ** 2 / (ref ** 2 + 1e-2)))


def random_rays(n, lo, hi, seed=0, tmax=np.inf):
    """Origins uniform in the box [lo, hi], uniformly random directions."""
    rng = np.random.default_rng(seed)
    rays = np.zeros(n, dtype=capi.RAY_DTYPE)
    rays["o"] = (rng.random((n, 3)) * (np.asarray(hi) - np.asarray(lo)) + np.asarray(lo)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    rays["d"] = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    rays["tmin"] = 1e-4
    rays["tmax"] = tmax
    return rays


def compare_hits(gpu, ref, second_t, min_bary, rays, rel_tol=1e-5):
    """Parity rule of BASELINE.json north_star (1): hit primitive bit-exact for non-degenerate rays, t within
    1e-5 relative.  Non-degenerate (SURVEY.md section 7 hard part iv): the oracle's hit lies more than 1e-6
    (barycentric) inside its triangle and no other triangle is hit within 1e-5 * t behind it.
    float32 cannot resolve t finer than the spacing of the coordinates it is computed from, so the bound on t
    is 1e-5 * t + 2 ulp(max(|o|_inf, t)); the second term only matters for hits much closer than |o|."""
    hit = np.isfinite(ref["t"])
    nondeg = ~hit | ((min_bary > 1e-6) & ~(second_t <= ref["t"] * (1 + 1e-5)))
    same_status = np.isfinite(gpu["t"]) == hit
    both = hit & np.isfinite(gpu["t"])
    same_prim = np.ones_like(hit)
    same_prim[both] = (gpu["prim"][both] == ref["prim"][both]) & (gpu["geom"][both] == ref["geom"][both])
    ok_t = np.ones_like(hit)
    scale = np.maximum(np.abs(rays["o"]).max(axis=1), np.where(hit, ref["t"], 0)).astype(np.float64)
    tol = rel_tol * np.abs(ref["t"][both].astype(np.float64)) + 2.0 * np.spacing(scale[both].astype(np.float32))
    ok_t[both] = np.abs(gpu["t"][both].astype(np.float64) - ref["t"][both]) <= tol
    pure_rel = np.ones_like(hit)
    pure_rel[both] = np.abs(gpu["t"][both].astype(np.float64) - ref["t"][both]) <= rel_tol * np.abs(ref["t"][both])
    bad = nondeg & ~(same_status & same_prim & ok_t)
    return dict(n=len(ref), nondegenerate=int(nondeg.sum()), hits=int(hit.sum()), mismatches=int(bad.sum()),
                prim_mismatches=int((nondeg & ~(same_status & same_prim)).sum()),
                t_outside_pure_relative=int((nondeg & ~pure_rel).sum()),
                bad_index=np.nonzero(bad)[0][:8], degenerate_disagreements=int((~nondeg & ~(same_status & same_prim)).sum()))
