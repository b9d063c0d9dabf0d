"""Summarise an .ncu-rep (ncu --set full) as JSON: per launch the duration, DRAM bytes, issue / tensor / DRAM utilisation.
  python tools/ncu_summary.py gpurun_out/x.ncu-rep "how it was captured" > profiles/x_summary.json"""
import csv
import json
import subprocess
import sys

rep, how = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
digest = sys.argv[3] if len(sys.argv) > 3 else None      # `python __graft_entry__.py --digest` on the box that captured
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}


def val(r, name, scale=1.0):
  if name not in col or r[col[name]] in ("", "n/a", "no data"):
    return None
  x = float(r[col[name]].replace(",", ""))
  u = units[col[name]]
  if u == "Gbyte":
    x *= 1e9
  elif u == "Mbyte":
    x *= 1e6
  elif u == "Kbyte":
    x *= 1e3
  elif u == "ms":
    x *= 1e6
  elif u == "us":
    x *= 1e3
  elif u == "s":
    x *= 1e9
  return x * scale


out = []
for r in rows[2:]:
  if len(r) < len(hdr):
    continue
  out.append({"kernel": r[col["Kernel Name"]], "grid": r[col.get("Grid Size", 0)], "block": r[col.get("Block Size", 0)],
              "duration_ms": val(r, "gpu__time_duration.sum", 1e-6),
              "dram_read_GB": val(r, "dram__bytes_read.sum", 1e-9), "dram_write_GB": val(r, "dram__bytes_write.sum", 1e-9),
              "dram_throughput_pct": val(r, "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
              "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
              "warps_active_pct": val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
              "inst_executed": val(r, "smsp__inst_executed.sum"),
              "tensor_pipe_active_pct": val(r, "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active"),
              "tensor_cycles_active_pct": val(r, "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
              "lts_hit_rate_pct": val(r, "lts__t_sector_hit_rate.pct"),
              "shared_bank_conflicts": val(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
              "registers_per_thread": val(r, "launch__registers_per_thread"),
              "sm_clock_ghz": val(r, "smsp__cycles_elapsed.avg.per_second",
                                  {"ghz": 1.0, "mhz": 1e-3, "hz": 1e-9}.get(units[col["smsp__cycles_elapsed.avg.per_second"]].lower(), 1.0)
                                  if "smsp__cycles_elapsed.avg.per_second" in col else 1.0)})
json.dump({"source": how, "csrc_digest": digest, "note": "per-launch values; ncu replays each kernel cold-cache and serialised: compare shares, "
           "not absolutes. GB = 1e9 bytes", "kernels": out}, sys.stdout, indent=1)
