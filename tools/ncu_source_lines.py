"""Stall samples of one profiled kernel by SOURCE LINE, for kernels whose code lives in inlined headers (the CUDA view
of `ncu --page source --csv` only lists the kernel's own file): joins the SASS view of the report with the line table of
the cubin (`nvdisasm -g`).

usage: python tools/ncu_source_lines.py REPORT.ncu-rep LAUNCH_INDEX OBJECT.o KERNEL_SUBSTRING [top]
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def line_table(obj, kernel):
    tmp = tempfile.mkdtemp()
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    text = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    table, cur, inside = {}, None, False
    for ln in text.splitlines():
        if ln.startswith(".text."):
            inside = kernel in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", ln)
        if m:
            table[int(m.group(1), 16)] = (cur, m.group(2).strip())
    return table


def main():
    rep, launch, obj, kernel = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip",
                          str(launch), "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    cols = {name: i for i, name in enumerate(rows[head])}
    table = line_table(obj, kernel)
    base = None
    by_line, by_reason = defaultdict(int), defaultdict(lambda: defaultdict(int))
    by_inst = defaultdict(int)
    total = total_inst = 0
    reasons = [c for c in cols if c.startswith("stall_") and "Not Issued" not in c]
    for r in rows[head + 1:]:
        if len(r) < len(cols) or r[0] == "Address":
            continue
        addr = int(r[cols["Address"]], 16)
        base = addr if base is None else base
        key, _ = table.get(addr - base, (None, ""))
        n = int(r[cols["# Samples"]] or 0)
        by_line[key] += n
        total += n
        ni = int(r[cols["Instructions Executed"]] or 0)
        by_inst[key] += ni
        total_inst += ni
        for c in reasons:
            by_reason[key][c] += int(r[cols[c]] or 0)
    print(f"# {rows[0][1][:100]}: {total} stall samples")
    for key, n in sorted(by_line.items(), key=lambda kv: -kv[1])[:top]:
        rs = sorted(by_reason[key].items(), key=lambda kv: -kv[1])[:3]
        print(f"{100.0 * n / total:6.2f}%  {key}  " + ", ".join(f"{a[6:]} {b}" for a, b in rs if b))
    print(f"# warp instructions executed by source line ({total_inst} in all)")
    for key, n in sorted(by_inst.items(), key=lambda kv: -kv[1])[:top]:
        print(f"{100.0 * n / total_inst:6.2f}%  {key}")


if __name__ == "__main__":
    main()
