"""Generic line-search routine — host mirror of src/cg_utils.jl."""
from __future__ import annotations


def evalϕdϕ_(xp, df_xp, fdf_, a, x, u):
    """evalϕdϕ! (src/cg_utils.jl:3-22).

    Reference: `xp[i] = x[i] + a*u[i]` (:13-15), `ϕ = fdf!(df_xp, xp)` (:18),
    `dϕ = dot(df_xp, u)` (:20) — three passes plus the callback.  Here: ONE fused kernel launch
    (`cgo_eval_trial`) that also leaves ‖df_xp‖² and every getβ dot in `xp.ws.pack`.
    `fdf_` is the device objective the workspace was built from (kept for signature parity).
    """
    ws = xp.ws
    assert ws is df_xp.ws is x.ws is u.ws
    return ws.eval_trial(a)
