"""Summarise an ncu report: key raw metrics per kernel + the top stall lines of the source page.
usage: ncu_top.py report.ncu-rep [kernel-regex] [n-lines]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else None
nl = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__issue_active.avg.pct",
        "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum", "sm__cycles_elapsed.max",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum", "smsp__inst_executed.sum"]
idx = [hdr.index(w) for w in want if w in hdr]
for r in rows[2:]:
    print({hdr[i]: (r[i][:48] + " " + units[i]) for i in idx})
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + (["--kernel-name", "regex:" + kre] if kre else []) +
                     ["--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
iS, iSrc, iEx = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
stalls = [i for i, x in enumerate(hdr) if x.startswith("stall_") and "Not Issued" not in x]
body = [r for r in rows[h + 1:] if len(r) > iS and r[iS].isdigit()]
tot = sum(int(r[iS]) for r in body)
print("total samples", tot)
for r in sorted(body, key=lambda r: -int(r[iS]))[:nl]:
    st = {hdr[i][6:]: int(r[i]) for i in stalls if r[i].isdigit() and int(r[i]) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(r[iS].rjust(6), r[iEx].rjust(9), r[iSrc].strip()[:64].ljust(64), st)
