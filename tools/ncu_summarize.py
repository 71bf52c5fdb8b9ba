#!/usr/bin/env python
"""Summaries of ncu outputs for profiles/ (run here, on files brought back in gpurun_out/).

  launches  <launches.csv>                 per-kernel count / total / average / share of an
                                            `ncu --metrics gpu__time_duration.sum --csv --log-file` launch list
  full      <report.ncu-rep>               one line per captured launch of a `--set full` report: duration, DRAM bytes,
                                            L2 hit rate, tensor-pipe activity, issue activity, registers, grid
  roles     <report.ncu-rep> <kernel regex> warp-stall samples of one kernel by SASS address range (roles of a
                                            warp-specialised kernel are contiguous address ranges): prints every mbarrier
                                            wait site with its barrier offset, the hottest instructions and a histogram
"""
import collections
import csv
import io
import re
import subprocess
import sys


def launches(path):
    rows = [r for r in csv.reader(open(path, newline="")) if len(r) > 10 and r[0].isdigit()]
    tot = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        name = re.sub(r"\(.*", "", r[4])
        tot[name][0] += 1
        tot[name][1] += float(r[-1]) / 1e3
    total = sum(v[1] for v in tot.values())
    for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:72]:72s} n={n:5d} total={us:10.1f}us avg={us / n:8.2f}us share={us / total:.3f}")
    print(f"total us {total:.1f}")


def _ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def full(rep):
    rows = _ncu_csv(rep, "raw")
    hdr = rows[0]
    want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size"]
    idx = [hdr.index(w) for w in want if w in hdr]
    print(" | ".join(hdr[i] for i in idx))
    print(" | ".join(rows[1][i] for i in idx))
    for r in rows[2:]:
        if len(r) == len(hdr):
            print(" | ".join(re.sub(r"\(.*", "", r[i]) if hdr[i] == "Kernel Name" else r[i] for i in idx))


def roles(rep, kernel):
    rows = _ncu_csv(rep, "source", ["--kernel-name", "regex:" + kernel])
    hdr = next(r for r in rows if r and r[0] == "Address")
    i_s, i_n = hdr.index("# Samples"), hdr.index("Instructions Executed")
    i0, i1 = hdr.index("stall_barrier"), hdr.index("stall_barrier (Not Issued)")
    data, seen = [], set()
    for r in rows:
        if not r or not r[0].startswith("0x") or len(r) != len(hdr) or r[0] in seen:
            continue
        seen.add(r[0])
        st = {hdr[i][6:]: int(r[i] or 0) for i in range(i0, i1)}
        data.append((int(r[0], 16), r[1].strip(), int(r[i_s]), int(r[i_n]), st))
    base, tot = data[0][0], sum(d[2] for d in data)
    print(f"{kernel}: {len(data)} SASS instructions, {tot} warp-stall samples")
    print("-- mbarrier wait sites (samples on the try_wait and the following 7 instructions)")
    for i, (a, ins, s, n, st) in enumerate(data):
        if "TRYWAIT" in ins:
            w = s + sum(d[2] for d in data[i + 1:i + 8] if "TRYWAIT" not in d[1])
            m = re.search(r"\+0x([0-9a-f]+)\]", ins)
            print(f"   idx {i:5d} +{a - base:6x} barrier@smem+0x{m.group(1) if m else '?'} executed {n:8d} samples {w:6d} "
                  f"{100 * w / tot:5.1f}%")
    print("-- hottest instructions")
    for i, (a, ins, s, n, st) in sorted(enumerate(data), key=lambda kv: -kv[1][2])[:25]:
        top = max(st.items(), key=lambda kv: kv[1])[0]
        print(f"   idx {i:5d} samples {s:6d} {100 * s / tot:5.1f}% executed {n:8d} {top:>16s}  {ins[:90]}")
    print("-- samples per 250-instruction window (roles are contiguous)")
    for lo in range(0, len(data), 250):
        seg = data[lo:lo + 250]
        s = sum(d[2] for d in seg)
        agg = collections.Counter()
        for d in seg:
            agg.update(d[4])
        print(f"   idx {lo:5d}-{lo + len(seg):5d} samples {s:6d} {100 * s / tot:5.1f}%  {agg.most_common(3)}")


if __name__ == "__main__":
    cmd = sys.argv[1]
    {"launches": launches, "full": full, "roles": roles}[cmd](*sys.argv[2:])
