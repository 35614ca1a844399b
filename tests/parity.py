"""One tolerance policy for every GPU parity assertion, and a log of what was measured.

POLICY (north star: "fp32 embeddings and gradients within rel 1e-5 / abs 1e-6", index work bit-exact)

  truth      fp64 runs: the reference in fp64 (`*_f64` fixtures) or the fp64 oracle.
             fp32 runs: the reference's fp64 arithmetic evaluated AT THE fp32-ROUNDED inputs and parameters (`*_r64` fixtures,
             or the oracle fed with the rounded tensors) - the exact value of the function at the point an fp32
             implementation is given.  (Rounding xi ~ 2K to fp32 alone moves the exact result by ~5e-6: input rounding,
             not kernel error.)
  STRICT     |got - truth| <= 1e-6 + 1e-5 |truth|  for EVERY element.  Used for all embedding outputs and for every
             gradient where it holds.
  SCALED     |got - truth| <= 1e-6 + 1e-5 |truth| + 1e-5 max|truth|.  Only for gradient tensors whose entries are sums of
             thousands of fp32 terms that cancel (dX, dtheta, dxi at K >= 100): an entry much smaller than its
             neighbours carries the rounding of terms as large as they are.  Every SCALED assertion also reports how many
             entries would fail STRICT, so the looser bar hides nothing.
  NEAR-STRICT  embedding VALUES at the benchmarked widths (K >= 127, frequencies up to xi ~ 2K) against the fp64-projection
             oracle: >= 99.95% of the entries STRICT and every entry within 1e-5 + 1e-5 |truth|
             (measured: 99.998% at K = 199, 99.978% at K = 511 where xi reaches 1021).  The projected keys are fp32
             in the reference (fsw_embedding.py:911) and here; the map keys -> out[., k] has gain (1+xi_k) sum_j |D_j| ~ 1.3 per
             element at xi w >> 1, so independent last-bit differences of the n keys of a segment move the top-frequency
             outputs by ~1.3 sqrt(n) ulp(p) ~ 2e-6 at n ~ 250 (measured: profiles/r2/README.md, all deviations sit in the last
             slices).  No implementation that keeps fp32 keys can be closer to the fp64-projection value; against the oracle
             evaluated ON THE KERNELS' OWN KEYS the same entries are STRICT (hubs of > 4000 elements: 2e-6).
  KEYS       gradients at those widths are asserted against the oracle evaluated on the kernels' own fp32 keys: dL/dp depends
             on the sorted ORDER, which is only defined up to ties ("bit-exact up to tie order"), and a projection summed in
             another order flips pairs of keys that agree to the last bit.  Rows holding an exactly equal key with another row
             are excluded and counted.
  fp64       rel 1e-9 / abs 1e-10 (gradients: abs 1e-9 max(1, max|truth|)).
  floor      when a fixture holds the reference's own fp32 result (`*_f32`), its deviation from the same truth is printed
             beside ours: the noise floor of the implementation we are a drop-in for.

Every call appends one line to LOG; tests/conftest.py prints the log in the terminal summary, so the driver's GPU test
record shows the measured deviations, not just PASSED.
"""
import numpy as np

RTOL, ATOL = 1e-5, 1e-6
LOG = []


def _stats(got, ref):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    err = np.abs(got - ref)
    lim = ATOL + RTOL * np.abs(ref)
    viol = err > lim
    mx = float(np.abs(ref).max()) if ref.size else 0.0
    return dict(n=int(ref.size), max_err=float(err.max()) if err.size else 0.0, max_ref=mx,
                strict_viol=int(viol.sum()), worst_ratio=float((err / lim).max()) if err.size else 0.0)


def check(name, got, ref, mode="strict", floor=None, fp64=False, grad=False):
    """Assert `got` against `ref` under the policy above and log the measurement.
    mode: 'strict' | 'scaled'.  floor: the reference's own fp32 result for the same quantity (reported only)."""
    got = np.asarray(got)
    ref = np.asarray(ref)
    assert got.shape == ref.shape, "%s: shape %s vs %s" % (name, got.shape, ref.shape)
    assert np.all(np.isfinite(got)), "%s: non-finite values" % name
    st = _stats(got, ref)
    line = "%-58s n=%-9d max|err|=%.2e (max|ref|=%.2e) strict-fail=%d worst err/lim=%.2f" % (
        name, st["n"], st["max_err"], st["max_ref"], st["strict_viol"], st["worst_ratio"])
    if floor is not None:
        fl = _stats(floor, ref)
        line += "  | reference-fp32 floor: max|err|=%.2e strict-fail=%d worst=%.2f" % (fl["max_err"], fl["strict_viol"], fl["worst_ratio"])
    if fp64:
        atol = 1e-9 * max(1.0, st["max_ref"]) if grad else 1e-10
        ok = np.abs(got.astype(np.float64) - ref.astype(np.float64)) <= atol + 1e-9 * np.abs(ref)
        line += "  [fp64 1e-9]"
    elif mode == "strict":
        ok = np.abs(got.astype(np.float64) - ref.astype(np.float64)) <= ATOL + RTOL * np.abs(ref)
        line += "  [STRICT]"
    elif mode == "near_strict":
        d = np.abs(got.astype(np.float64) - ref.astype(np.float64))
        frac = float((d > ATOL + RTOL * np.abs(ref)).mean()) if ref.size else 0.0
        ok = np.array([frac <= 5e-4, bool(np.all(d <= 1e-5 + RTOL * np.abs(ref)))])
        line += "  [NEAR-STRICT: %.5f%% beyond strict]" % (100 * frac)
    elif mode == "scaled":
        ok = np.abs(got.astype(np.float64) - ref.astype(np.float64)) <= ATOL + RTOL * np.abs(ref) + RTOL * st["max_ref"]
        line += "  [SCALED]"
    else:
        raise ValueError(mode)
    line += " OK" if bool(np.all(ok)) else " FAIL"
    LOG.append(line)
    assert bool(np.all(ok)), line
    return st
