"""Registers / spills per kernel from the ptxas logs the build keeps (csrc/build/*.ptxas.txt).  usage: regs.py [regex]"""
import glob, re, subprocess, sys
pat = re.compile(sys.argv[1]) if len(sys.argv) > 1 else None
for f in sorted(glob.glob("rl_ctr_prediction_b200/csrc/build/*.ptxas.txt")):
    name = None
    for line in open(f):
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            name = m.group(1)
        m = re.search(r"Used (\d+) registers", line)
        if m and name:
            dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip().split("(")[0]
            if pat is None or pat.search(dem):
                print(f"{m.group(1):>4} {dem}")
            name = None
