"""Comparison helpers shared by the parity tests."""
import numpy as np

# what a closed-loop case records, by exactness class
INDEX_KEYS = ("best", "steps", "reached", "known", "best_type", "best_id")
FLOAT_KEYS = ("next_pos", "next_vel", "length", "min_obs_dist", "goal_dist", "final_paths", "final_vel",
              "trajectory", "rot")


def assert_bit_identical(got, want, keys=None, ctx=""):
    for k in keys or want.keys():
        if k == "sha_inputs" or k not in got:
            continue
        a, b = np.asarray(got[k]), np.asarray(want[k])
        assert a.shape == b.shape, f"{ctx}{k}: shape {a.shape} != {b.shape}"
        if b.dtype.kind == "f":
            # bit patterns, so +0.0 and -0.0 differ; any NaN matches any NaN (payloads are not part of the contract)
            a, b = np.ascontiguousarray(a, dtype=np.float64), np.ascontiguousarray(b, dtype=np.float64)
            same = (a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b))
        else:
            same = a == b
        if not np.all(same):
            idx = np.argwhere(~same)
            first = tuple(idx[0])
            raise AssertionError(f"{ctx}{k}: {len(idx)} of {a.size} differ; first at {first}: got {a[first]!r} want {b[first]!r}")


def assert_close(got, want, rtol, keys=FLOAT_KEYS, ctx=""):
    """north_star tolerance: |got - want| <= rtol * max(|want|, 1) elementwise, NaNs must coincide."""
    for k in keys:
        if k not in got or k not in want:
            continue
        a, b = np.asarray(got[k], dtype=np.float64), np.asarray(want[k], dtype=np.float64)
        assert a.shape == b.shape, f"{ctx}{k}: shape {a.shape} != {b.shape}"
        nan_a, nan_b = np.isnan(a), np.isnan(b)
        assert np.array_equal(nan_a, nan_b), f"{ctx}{k}: NaN pattern differs"
        err = np.abs(np.where(nan_a, 0.0, a - b))
        tol = rtol * np.maximum(np.abs(np.where(nan_b, 0.0, b)), 1.0)
        if not np.all(err <= tol):
            i = np.unravel_index(np.argmax(err - tol), err.shape)
            raise AssertionError(f"{ctx}{k}: |err|={err[i]:.3e} > tol={tol[i]:.3e} at {i} (got {a[i]!r}, want {b[i]!r})")
