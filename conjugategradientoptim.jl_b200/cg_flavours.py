"""CG direction plugins — host mirror of src/cg_flavours.jl (scalar logic only).

Every vector pass of the reference's getβ (3 temporaries, 3–6 dots, SURVEY.md §8a B1–B4) was
already done by the trial kernel; getβ here combines the dot pack on the host so that Julia's
NaN / Inf / `max` semantics are kept exactly.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from ._capi import P_DPHI, P_GPG, P_GPGP, P_UG, P_UU, P_UY, P_YGP, P_YY
from .cg_types import CGβConfig

f64 = np.float64


# ---- generic core routines
def updatedir_(u, df_x, β):
    """updatedir! (cg_flavours.jl:2-15): u[i] = −df_x[i] + β*u[i]."""
    assert u.ws is df_x.ws                                          # :8
    if isinstance(β, LBFGSHistoryToken):                            # quasi-Newton flavour
        u.ws.lbfgs_update_dir()
    else:
        u.ws.update_dir(β)


def initializeβ(β_config):
    """cg_flavours.jl:17-19"""
    return f64(0.0)


def initializeLineSearchContainer_(info, β_config, df_x, x):
    """cg_flavours.jl:22-35: u = −df_x; x, xp ← x; df_xp ← df_x (the device state already holds
    x and df_x; only the direction needs a kernel)."""
    info.reset_direction()


class LBFGSHistoryToken:
    """What getβ returns for the LBFGS flavour (the history itself lives on the device)."""


def _jl_max(a, b):
    """Julia `max`: NaN-propagating (cg_flavours.jl:68)."""
    if a != a:
        return a
    if b != b:
        return b
    return a if a > b else b


# ---- (Yuan 2019): modified Hager-Zhang with trust-region behaviour
@dataclass(frozen=True)
class YuanWangSheng(CGβConfig):
    μ: float                                                        # cg_flavours.jl:46-48


@dataclass(frozen=True)
class HagerZhang(CGβConfig):                                        # cg_flavours.jl:83
    pass


@dataclass(frozen=True)
class SallehAlhawarat(CGβConfig):                                   # cg_flavours.jl:130
    pass


@dataclass(frozen=True)
class LiuStorrey(CGβConfig):                                        # cg_flavours.jl:154
    pass


def _hz_family(ws, R):
    P = ws.pack
    m = 2 * f64(P[P_YY]) / R                                        # :73 / :102
    if ws.beta_form == "literal":
        # tmp2 = g_next ./ R; tmp1 = y − m .* u; dot(tmp1, tmp2)     :71-76 / :100-105
        return ws.beta_literal(R, m)
    # single-pass form: Σ (y_i − m u_i)(g⁺_i / R) = (y·g⁺ − m u·g⁺) / R
    return (f64(P[P_YGP]) - m * f64(P[P_DPHI])) / R


def getβ(β_config, g_next, g, u):
    """getβ (cg_flavours.jl:51-79, 87-108, 133-151, 157-170), dispatched on the flavour type.
    g_next = info.df_xp, g = df_x, u = info.u: the dots below were reduced by the trial kernel
    that produced g_next (y = g_next − g formed elementwise there, as the reference does)."""
    ws = g_next.ws
    P = ws.pack
    with np.errstate(all="ignore"):
        if isinstance(β_config, YuanWangSheng):
            μ = f64(β_config.μ)
            R1 = μ * np.sqrt(f64(P[P_UU])) * np.sqrt(f64(P[P_YY]))      # :65
            R2 = f64(P[P_UY])                                           # :66
            R3 = 2 * f64(P[P_YY]) * f64(P[P_DPHI]) / f64(P[P_YGP])      # :67
            R = _jl_max(_jl_max(R1, R2), R3)                            # :68
            return _hz_family(ws, R)
        if isinstance(β_config, HagerZhang):
            R = f64(P[P_UY])                                            # :98
            return _hz_family(ws, R)
        if isinstance(β_config, SallehAlhawarat):
            nrm = np.sqrt(f64(P[P_GPGP]))
            norm_sq = nrm * nrm                                         # :140 norm(g_next)^2
            tmp = f64(P[P_GPG])                                         # :141
            if norm_sq > tmp:
                numerator = norm_sq - tmp
                denominator = f64(P[P_DPHI]) - f64(P[P_UG])             # :145
                return numerator / denominator
            return f64(0.0)
        if isinstance(β_config, LiuStorrey):
            numerator = f64(P[P_YGP])                                   # :166
            denominator = -f64(P[P_UY])                                 # :167
            return numerator / denominator
    from .qn_flavours import LBFGS, BroydenFamily, getβ_lbfgs
    if isinstance(β_config, LBFGS):
        return getβ_lbfgs(β_config, ws)
    if isinstance(β_config, BroydenFamily):
        return f64(0.0)             # B never leaves the identity (qn_flavours.jl:81): u = −g
    raise TypeError(f"no getβ method for {type(β_config).__name__}")
