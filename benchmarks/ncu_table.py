"""ncu report(s) -> one markdown table row per kernel launch: time, DRAM bytes, achieved GB/s, L2 / DRAM / SM throughput
percentages, occupancy, top three stall reasons.  usage: ncu_table.py REPORT.ncu-rep [REPORT2 ...] > profiles/x.md"""
import csv
import subprocess
import sys

STALLS = ["long_scoreboard", "short_scoreboard", "barrier", "wait", "lg_throttle", "mio_throttle", "math_pipe_throttle",
          "not_selected", "branch_resolving", "no_instruction", "membar", "dispatch_stall", "tex_throttle", "sleeping",
          "drain", "imc_miss"]


def num(x):
    try:
        return float(x.replace(",", ""))
    except (ValueError, AttributeError):
        return None


print("| kernel | grid x block | regs | time (us) | DRAM rd+wr (MB) | DRAM GB/s | dram % | L2 % | SM busy % | occupancy % | "
      "issue active % | top stalls (warps per issue) |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|")
for path in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        continue
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(h, r))
        u = dict(zip(h, units))
        t = num(d.get("gpu__time_duration.sum"))
        if t is not None and u.get("gpu__time_duration.sum", "").startswith("ns"):
            t /= 1e3
        elif t is not None and u.get("gpu__time_duration.sum", "").startswith("ms"):
            t *= 1e3
        elif t is not None and u.get("gpu__time_duration.sum", "") in ("s", "second"):
            t *= 1e6

        def bytes_of(key):
            v, unit = num(d.get(key)), u.get(key, "")
            if v is None:
                return 0.0
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)
        dram = bytes_of("dram__bytes_read.sum") + bytes_of("dram__bytes_write.sum")
        stalls = []
        for s in STALLS:
            v = num(d.get("smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s))
            if v is not None:
                stalls.append((v, s))
        stalls.sort(reverse=True)
        name = d.get("Kernel Name", "?").split("(")[0][-60:]
        print("| `%s` | %s x %s | %s | %.1f | %.2f | %.0f | %s | %s | %s | %s | %s | %s |" % (
            name, d.get("launch__grid_size", "?"), d.get("launch__block_size", "?"), d.get("launch__registers_per_thread", "?"),
            t or 0.0, dram / 1e6, dram / (t * 1e-6) / 1e9 if t else 0.0,
            d.get("dram__throughput.avg.pct_of_peak_sustained_elapsed", "?"),
            d.get("lts__throughput.avg.pct_of_peak_sustained_elapsed", "?"),
            d.get("sm__throughput.avg.pct_of_peak_sustained_elapsed", "?"),
            d.get("sm__warps_active.avg.pct_of_peak_sustained_active", "?"),
            d.get("smsp__issue_active.avg.pct_of_peak_sustained_active", "?"),
            ", ".join("%s %.2f" % (s, v) for v, s in stalls[:3])))
