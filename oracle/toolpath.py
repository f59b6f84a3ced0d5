"""Oracle (test infrastructure): G-code -> fixed-width toolpath rows.

Restates parsingGcode cP:6-188, count_lines cP:191-196, format_fixed cP:199-201.  A row is
``x,y,z,Ljump,Ldwell,dt,P`` (cP:71-74); x,y,z are right-justified ``%.8e`` in 15 columns so
every row has the same byte length (the checkpoint ``seek`` at gm:125-126 depends on it).
Checked byte-for-byte against the reference's own parser (tests/golden/toolpath_example.txt,
made by tests/golden/make_golden.py running cP:6-188 unmodified).
"""
import re

import numpy as np

_G_RE = re.compile(
    r"(?:G(\d+)\s*X(-?\d+\.\d+|-?\d+)\s*Y(-?\d+\.\d+|-?\d+)(?:\s*Z(-?\d+\.\d+|-?\d+))?)"
)


def format_fixed(val, width=15, precision=8):
    return f"{val:.{precision}e}".rjust(width)


def _row(x, y, z, jump, dwell, dt, P):
    return (
        f"{format_fixed(x)},{format_fixed(y)},{format_fixed(z)},"
        f"{jump:d},{dwell:d},{dt:.8e},{P:.8e}\n"
    )


def _waypoints(gcode_text):
    """cP:28-48: (x, y, z, is_jump_target) per G-line; z sticks from the last line that gave one."""
    pts = []
    z = None
    for cmd, sx, sy, sz in _G_RE.findall(gcode_text):
        if sz:
            z = float(sz)
        pts.append((float(sx), float(sy), z, 0.0 if cmd == "1" else 1.0))
    return pts


def _dwell_rows(x, y, z, jump, nm, dwell_t, coef):
    """cP:66-103 and cP:162-187: wait_time small steps, then coarse dwell steps, then remainder."""
    dt3 = nm["timestep_L3"]
    rows = []
    for k in range(int(nm["wait_time"])):
        if (k + 1) * dt3 > nm["dwell_time"]:
            break
        rows.append(_row(x, y, z, jump, 0, dt3, 0))
    rest = max(0, dwell_t - nm["wait_time"] * dt3)
    nbig = int(rest / dt3 / coef)
    for _ in range(nbig):
        rows.append(_row(x, y, z, jump, 0, dt3 * coef, 0))
    small = rest - nbig * dt3 * coef
    if small > 0:
        rows.append(_row(x, y, z, jump, 0, small, 0))
    return rows


def toolpath_rows(Nonmesh, Properties):
    nm = Nonmesh
    coef = float(nm["dwell_time_multiplier"] * nm["subcycle_num_L2"] * nm["subcycle_num_L3"])
    dwell_t = max(nm["dwell_time"] - nm["wait_time"] * nm["timestep_L3"], nm["dwell_time"])
    with open(nm["gcode"], "r") as fh:
        pts = _waypoints(fh.read())
    step_len = nm["laser_velocity"] * nm["timestep_L3"]
    rows = []
    z = pts[0][2]
    x = y = None
    jump = 1
    for a, b in zip(pts[:-1], pts[1:]):
        jump = 1
        if a[2] != b[2]:  # layer change: dwell at the last point of the old layer (cP:61-104)
            x, y = a[0], a[1]
            if dwell_t > 0:
                rows += _dwell_rows(x, y, z, jump, nm, dwell_t, coef)
            else:
                for k in range(int(nm["wait_time"])):
                    if (k + 1) * nm["timestep_L3"] > nm["dwell_time"]:
                        break
                    rows.append(_row(x, y, z, jump, 0, nm["timestep_L3"], 0))
            continue
        if b[3] == 1:
            jump = 0
        seg = np.linalg.norm(np.array([b[0] - a[0], b[1] - a[1]]))
        nfull = int(seg // step_len)
        shortdt = (seg % step_len) / step_len * nm["timestep_L3"]
        x, y = a[0], a[1]
        vx = nm["laser_velocity"] * (b[0] - a[0]) / seg
        vy = nm["laser_velocity"] * (b[1] - a[1]) / seg
        z = a[2]
        P = jump * Properties["laser_power"]
        for _ in range(nfull):
            x += vx * nm["timestep_L3"]
            y += vy * nm["timestep_L3"]
            rows.append(_row(x, y, z, jump, 1, nm["timestep_L3"], P))
        if shortdt > 0:
            x += vx * shortdt
            y += vy * shortdt
            rows.append(_row(x, y, z, jump, 1, shortdt, P))
    # cP:158-187: final dwell (no dwell_t > 0 guard here in the reference)
    rows += _dwell_rows(x, y, z, jump, nm, dwell_t, coef)
    return rows


def parsingGcode(Nonmesh, Properties, L2h=None):
    """cP:6-188: write Nonmesh['toolpath'], return the number of rows."""
    rows = toolpath_rows(Nonmesh, Properties)
    with open(Nonmesh["toolpath"], "w") as out:
        out.writelines(rows)
    return len(rows)


def count_lines(file_path):
    """cP:191-196."""
    with open(file_path, "r") as fh:
        return sum(1 for _ in fh)
