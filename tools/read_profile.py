#!/usr/bin/env python
"""Reader for the files written with UCC_PROFILE_MODE=log|accum (role of `ucx_read_profile` for the reference).

    python tools/read_profile.py ucc_host_1234.prof                # accumulated table + per-request latencies
    python tools/read_profile.py ucc_host_1234.prof --chrome t.json  # chrome://tracing / Perfetto timeline
"""
from __future__ import annotations

import argparse
import json
import sys
from collections import defaultdict


def parse(path):
    accum, log = [], []
    with open(path) as f:
        for ln in f:
            p = ln.split()
            if not p or p[0].startswith("#"):
                continue
            if p[0] == "A" and len(p) >= 6:
                accum.append({"name": p[1], "where": p[2], "count": int(p[3]), "total_us": float(p[4]), "avg_us": float(p[5])})
            elif p[0] == "L" and len(p) >= 5:
                log.append({"name": p[1], "t": float(p[2]), "dur_us": float(p[3]), "req": p[4]})
    return accum, log


def request_latencies(log):
    """pair <x>_start / <x>_done (and ucc_collective_post / ucc_coll_complete) events of the same request"""
    open_ev, out = {}, defaultdict(list)
    for e in log:
        n = e["name"]
        if n.endswith("_start"):
            open_ev[(n[:-6], e["req"])] = e["t"]
        elif n.endswith("_done"):
            t0 = open_ev.pop((n[:-5], e["req"]), None)
            if t0 is not None:
                out[n[:-5]].append((e["t"] - t0) * 1e6)
    return out


def chrome_trace(log):
    t_base = min((e["t"] for e in log), default=0.0)
    ev = []
    for e in log:
        ts = (e["t"] - t_base) * 1e6
        if e["dur_us"] > 0:
            ev.append({"name": e["name"], "ph": "X", "ts": ts, "dur": e["dur_us"], "pid": 0, "tid": 0})
        else:
            ev.append({"name": e["name"], "ph": "i", "ts": ts, "s": "t", "pid": 0, "tid": 1, "args": {"req": e["req"]}})
    return {"traceEvents": ev, "displayTimeUnit": "ns"}


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("file")
    ap.add_argument("--chrome", help="write a chrome-trace JSON of the logged records here")
    ap.add_argument("--json", action="store_true", help="machine readable output")
    a = ap.parse_args(argv)
    accum, log = parse(a.file)
    lat = request_latencies(log)
    if a.chrome:
        with open(a.chrome, "w") as f:
            json.dump(chrome_trace(log), f)
    if a.json:
        print(json.dumps({"accum": accum, "n_log": len(log),
                          "latency_us": {k: {"n": len(v), "avg": sum(v) / len(v), "min": min(v), "max": max(v)} for k, v in lat.items()}}))
        return 0
    if accum:
        print(f"{'location':36s} {'count':>8s} {'total us':>12s} {'avg us':>10s}  where")
        for r in sorted(accum, key=lambda r: -r["total_us"]):
            print(f"{r['name']:36s} {r['count']:8d} {r['total_us']:12.1f} {r['avg_us']:10.3f}  {r['where']}")
    if lat:
        print(f"\n{'request (start -> done)':36s} {'n':>8s} {'avg us':>10s} {'min us':>10s} {'max us':>10s}")
        for k, v in sorted(lat.items()):
            print(f"{k:36s} {len(v):8d} {sum(v) / len(v):10.2f} {min(v):10.2f} {max(v):10.2f}")
    if not accum and not log:
        print("no records", file=sys.stderr)
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
