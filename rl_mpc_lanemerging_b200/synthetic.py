"""Seeded synthetic on-ramp traffic states (SURVEY.md §8(d) "Synthetic inputs").

There is no SUMO in the container, so benchmark/test inputs are generated: highway cars are laid
out the way the reference's spawner spaces them (control.py:215-226: one car every
BASE_TRAFFIC_INTERVAL + U[0,1) seconds at OTHER_CAR_SPEED), the ego sits on the ramp with the
reference's start-speed distribution (control.py:198-204), and only cars inside SENSOR_RADIUS are
kept, sorted front->back (prediction.py:134-141).

The generator is counter based: every random number is a hash of (seed, global episode id, draw
index), so the state of episode e does not depend on the batch size or on how episodes are sharded
over GPUs.
"""
from __future__ import annotations

import numpy as np

# (BASE_TRAFFIC_INTERVAL [s], OTHER_CAR_SPEED [m/s]) of configs/st_*.json
TRAFFIC = {"low": (2.4, 7.0), "medium": (1.8, 7.0), "default": (1.2, 7.0), "moderate": (1.2, 11.0),
           "fast": (1.2, 15.0)}
NMAX = 32
SENSOR_RADIUS = 125.0
# ramp lane centre line, linear approximation of merge.net.xml:52
RAMP_A = (-250.47, 28.47)
RAMP_B = (-50.58, 1.71)
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return x ^ (x >> np.uint64(31))


def uniform(seed: int, episode: np.ndarray, draw: np.ndarray | int) -> np.ndarray:
    """U[0,1) as a pure function of (seed, episode, draw)."""
    with np.errstate(over="ignore"):
        e = episode.astype(np.uint64)
        d = np.asarray(draw).astype(np.uint64)
        k = _splitmix(np.uint64(seed) + np.uint64(0x632BE59BD9B4E019))
        h = _splitmix(_splitmix(e ^ k) + d * np.uint64(0xD1B54A32D192ED03))
    return (h >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def make_states(batch: int, traffic: str = "moderate", seed: int = 0, first_episode: int = 0,
                kind: str = "onramp", nmax: int = NMAX):
    """Returns dict(ego[B,4]=(x,y,v,a) f64, cars_x/v/a[B,nmax] f64, n_cars[B] i32).

    kind="onramp": ego on the ramp, x~U(-210,-60), v=clip(N(15,5),5,25), a~U(-1,1); constant-speed traffic.
    kind="mixed":  ego anywhere on ramp or highway (x~U(-210,70)); car speeds/accelerations perturbed so
                   that the follower model, merged-ego branches and infeasible plans are exercised.
    """
    interval, speed = TRAFFIC[traffic]
    ep = np.arange(first_episode, first_episode + batch, dtype=np.int64)
    u = lambda d: uniform(seed, ep, d)  # noqa: E731
    # --- ego
    if kind == "onramp":
        ex = -210.0 + 150.0 * u(0)
    elif kind == "mixed":
        ex = -210.0 + 280.0 * u(0)
    else:
        raise ValueError(kind)
    frac = (ex - RAMP_A[0]) / (RAMP_B[0] - RAMP_A[0])
    ey = RAMP_A[1] + frac * (RAMP_B[1] - RAMP_A[1])
    if kind == "mixed":
        # past the end of the ramp polyline the ego heads for / sits in the highway lane (y=-1.6)
        t = np.clip((ex - RAMP_B[0]) / (1.5 - RAMP_B[0]), 0.0, 1.0)
        ey = np.where(ex > RAMP_B[0], RAMP_B[1] + t * (-1.6 - RAMP_B[1]), ey)
    z = np.sqrt(-2.0 * np.log(1.0 - u(1))) * np.cos(2.0 * np.pi * u(2))
    ev = np.clip(15.0 + 5.0 * z, 5.0, 25.0)
    ea = -1.0 + 2.0 * u(3)
    # --- highway cars, front -> back
    ncand = 64
    d = np.arange(ncand)[None, :]
    epc = ep[:, None]
    gaps = speed * (interval + uniform(seed, epc, 16 + d))
    gaps[:, 0] = uniform(seed, ep, 16) * speed * (interval + 0.5)
    xs = 100.0 - np.cumsum(gaps, axis=1)
    vs = np.full_like(xs, speed)
    acs = np.zeros_like(xs)
    if kind == "mixed":
        vs = vs * (0.6 + 0.6 * uniform(seed, epc, 128 + d))
        acs = np.where(uniform(seed, epc, 256 + d) < 0.3, -4.0 * uniform(seed, epc, 384 + d), 0.0)
    dist = np.sqrt((xs - ex[:, None]) ** 2 + (-1.6 - ey[:, None]) ** 2)
    keep = (xs >= -250.0) & (dist < SENSOR_RADIUS)
    order = np.argsort(~keep, axis=1, kind="stable")      # kept cars first, front->back order preserved
    xs = np.take_along_axis(xs, order, 1)[:, :nmax]
    vs = np.take_along_axis(vs, order, 1)[:, :nmax]
    acs = np.take_along_axis(acs, order, 1)[:, :nmax]
    n = np.minimum(keep.sum(1), nmax).astype(np.int32)
    pad = np.arange(nmax)[None, :] >= n[:, None]
    xs[pad] = 0.0
    vs[pad] = 0.0
    acs[pad] = 0.0
    ego = np.stack([ex, ey, ev, ea], 1)
    return dict(ego=np.ascontiguousarray(ego), cars_x=np.ascontiguousarray(xs), cars_v=np.ascontiguousarray(vs),
                cars_a=np.ascontiguousarray(acs), n_cars=n)


def horizon_settings(H: int):
    """FUTURE_T / FUTURE_S that give exactly H+1 time layers and keep FUTURE_S = MAX_SPEED*FUTURE_T
    (SURVEY.md §8(d)); H=17 is the published configuration (FUTURE_T 5.0, FUTURE_S 150)."""
    if H == 17:
        return 5.0, 150.0
    return 0.3 * H - 0.15, 30.0 * 0.3 * H
