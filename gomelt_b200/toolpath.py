"""G-code -> toolpath text of the drop-in (host side; parsingGcode cP:6-188, count_lines cP:191-196).

One text row per time step: ``x,y,z,Ljump,Ldwell,dt,P`` with x, y, z as right-justified ``%.8e`` in 15
columns, so every row has the same byte length - the checkpoint restart seeks by row
(gm:125-126).  Implemented as a generator of row tuples (``iter_rows``) + one formatter; the file is
what ``go_melt.py`` reads back N2*N3 rows at a time (gm:140-143).
"""
import re

import numpy as np

_LINE = re.compile(r"G(\d+)\s*X(-?\d+(?:\.\d+)?)\s*Y(-?\d+(?:\.\d+)?)(?:\s*Z(-?\d+(?:\.\d+)?))?")


def _fmt(row):
    x, y, z, jump, on, dt, power = row
    return "{},{},{},{:d},{:d},{:.8e},{:.8e}\n".format(*(f"{c:.8e}".rjust(15) for c in (x, y, z)), jump, on, dt, power)


def read_waypoints(path):
    """[(x, y, z, is_rapid)]: G1 = scan move (laser on), anything else = rapid; z persists (cP:28-48)."""
    pts, z = [], None
    with open(path, "r") as fh:
        for cmd, sx, sy, sz in _LINE.findall(fh.read()):
            z = float(sz) if sz else z
            pts.append((float(sx), float(sy), z, cmd != "1"))
    return pts


def _pause(x, y, z, jump, nm):
    """Laser-off rows at a fixed position: up to ``wait_time`` fine steps, then the remaining dwell in
    coarse steps of dwell_time_multiplier*N2*N3 fine steps, then what is left (cP:66-103, 162-187)."""
    dt = nm["timestep_L3"]
    coef = float(nm["dwell_time_multiplier"] * nm["subcycle_num_L2"] * nm["subcycle_num_L3"])
    big = dt * coef
    for i in range(1, int(nm["wait_time"]) + 1):
        if i * dt > nm["dwell_time"]:
            break
        yield (x, y, z, jump, 0, dt, 0)
    # the reference's max(dwell - wait*dt, dwell) is just dwell (cP:14-17)
    remaining = max(0, max(nm["dwell_time"] - nm["wait_time"] * dt, nm["dwell_time"]) - nm["wait_time"] * dt)
    n_big = int(remaining / dt / coef)  # the reference's operation order (cP:82-84, 173): it decides the row count
    for _ in range(n_big):
        yield (x, y, z, jump, 0, big, 0)
    tail = remaining - n_big * dt * coef
    if tail > 0:
        yield (x, y, z, jump, 0, tail, 0)


def iter_rows(nonmesh, properties):
    nm = nonmesh
    dt, vel = nm["timestep_L3"], nm["laser_velocity"]
    stride = vel * dt
    pts = read_waypoints(nm["gcode"])
    z = pts[0][2]
    x = y = None
    jump = 1
    dwell_positive = max(nm["dwell_time"] - nm["wait_time"] * dt, nm["dwell_time"]) > 0
    for (ax, ay, az, _), (bx, by, bz, rapid) in zip(pts, pts[1:]):
        jump = 1
        if az != bz:  # new layer: pause where the old layer ended (cP:61-104)
            x, y = ax, ay
            if dwell_positive:
                yield from _pause(x, y, z, jump, nm)
            else:
                for i in range(1, int(nm["wait_time"]) + 1):
                    if i * dt > nm["dwell_time"]:
                        break
                    yield (x, y, z, jump, 0, dt, 0)
            continue
        if rapid:
            jump = 0
        dist = float(np.linalg.norm(np.array([bx - ax, by - ay])))
        whole = int(dist // stride)
        rest_dt = (dist % stride) / stride * dt
        ux, uy = vel * (bx - ax) / dist, vel * (by - ay) / dist
        x, y, z = ax, ay, az
        power = jump * properties["laser_power"]
        for _ in range(whole):
            x += ux * dt
            y += uy * dt
            yield (x, y, z, jump, 1, dt, power)
        if rest_dt > 0:
            x += ux * rest_dt
            y += uy * rest_dt
            yield (x, y, z, jump, 1, rest_dt, power)
    yield from _pause(x, y, z, jump, nm)  # final dwell (cP:158-187)


def parsingGcode(Nonmesh, Properties, L2h=None):
    """Write ``Nonmesh['toolpath']``; returns the number of rows (cP:6-188)."""
    n = 0
    with open(Nonmesh["toolpath"], "w") as out:
        for row in iter_rows(Nonmesh, Properties):
            out.write(_fmt(row))
            n += 1
    return n


def count_lines(file_path):
    with open(file_path, "r") as fh:
        return sum(1 for _ in fh)
