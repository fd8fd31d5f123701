"""Per-source-line view of one kernel of an ncu report (captured with --import-source on, built with -lineinfo).

    python tools/ncu_by_line.py gpurun_out/x.ncu-rep aspect_b200/csrc/operators.o 'velocity_kernel<3, 0' [out.txt]

Joins the SASS page of the report (instructions executed, stall samples, shared-memory wavefronts per instruction)
with nvdisasm's line table of the object file and sums per CUDA source line.
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def sass_rows(rep, kernel_regex):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name-base", "demangled", "--kernel-name", "regex:" + kernel_regex,
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    name = rows[0][1]
    hdr = rows[1]
    body = []
    for r in rows[2:]:
        if r and r[0] == "Kernel Name":   # next kernel of the report that matches the regex: keep the first only
            break
        body.append(r)
    return name, hdr, body


def line_table(obj, mangled_part):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "--print-line-info", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    table, cur, inside = {}, None, False
    for ln in dis.splitlines():
        if ln.startswith(".text."):
            inside = mangled_part in ln
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*);", ln)
        if m:
            table[int(m.group(1), 16)] = cur
    return table


def main():
    rep, obj, kre = sys.argv[1:4]
    name, hdr, rows = sass_rows(rep, kre)
    # mangled name fragment: template args of the demangled name -> e.g. velocity_kernelILi3ELi0ELb0ELb1E
    m = re.search(r"(\w+)<\(int\)(\d+), \(int\)(\d+), \(bool\)(\d+), \(bool\)(\d+)>", name)
    if m:
        frag = f"{m.group(1)}ILi{m.group(2)}ELi{m.group(3)}ELb{m.group(4)}ELb{m.group(5)}E"
    else:
        m = re.search(r"(\w+)<\(int\)(\d+), \(bool\)(\d+)>", name)
        frag = f"{m.group(1)}ILi{m.group(2)}ELb{m.group(3)}E" if m else re.search(r"(\w+)[<(]", name).group(1)
    table = line_table(obj, frag)
    col = {k: hdr.index(k) for k in ["Address", "Source", "# Samples", "Instructions Executed", "L1 Wavefronts Shared",
                                      "L1 Wavefronts Shared Excessive", "stall_long_sb", "stall_short_sb", "stall_wait",
                                      "stall_barrier", "stall_math", "stall_mio", "stall_lg", "L2 Theoretical Sectors Global"]}
    agg = collections.defaultdict(lambda: collections.Counter())
    base = None
    tot = collections.Counter()
    for r in rows:
        if len(r) < len(hdr):
            continue
        addr = int(r[col["Address"]], 16) if r[col["Address"]].startswith("0x") else int(r[col["Address"]])
        if base is None:
            base = addr
        key = table.get(addr - base, ("?", 0))
        a = agg[key]
        for k in col:
            if k in ("Address", "Source"):
                continue
            try:
                v = float(r[col[k]] or 0)
            except ValueError:
                v = 0.0
            a[k] += v; tot[k] += v
        op = r[col["Source"]].split()[0] if r[col["Source"]] else ""
        if op.startswith("@"):
            op = r[col["Source"]].split()[1]
        a["op:" + op.split(".")[0]] += float(r[col["Instructions Executed"]] or 0)
    src = {}
    lines = []
    lines.append(f"# {name}\n# totals: " + ", ".join(f"{k}={v:.0f}" for k, v in tot.items()))
    lines.append(f"{'line':>14s} {'inst':>10s} {'%inst':>6s} {'samples':>8s} {'%smp':>6s} {'long_sb':>8s} {'short_sb':>8s} {'wait':>7s} {'barrier':>7s} "
                 f"{'smem_wf':>10s} {'smem_exc':>10s} {'gl_sect':>10s}  ops")
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"]):
        f, l = key
        if f not in src and f != "?":
            p = os.path.join(os.path.dirname(obj), f)
            src[f] = open(p).read().splitlines() if os.path.exists(p) else []
        text = src.get(f, [])[l - 1].strip()[:90] if f != "?" and l - 1 < len(src.get(f, [])) else ""
        ops = " ".join(f"{k[3:]}:{v:.0f}" for k, v in sorted(a.items(), key=lambda kv: -kv[1]) if k.startswith("op:"))[:80]
        lines.append(f"{f[:8]:>8s}:{l:<5d} {a['Instructions Executed']:10.0f} {100 * a['Instructions Executed'] / max(tot['Instructions Executed'], 1):6.2f} "
                     f"{a['# Samples']:8.0f} {100 * a['# Samples'] / max(tot['# Samples'], 1):6.2f} {a['stall_long_sb']:8.0f} {a['stall_short_sb']:8.0f} "
                     f"{a['stall_wait']:7.0f} {a['stall_barrier']:7.0f} {a['L1 Wavefronts Shared']:10.0f} {a['L1 Wavefronts Shared Excessive']:10.0f} "
                     f"{a['L2 Theoretical Sectors Global']:10.0f}  | {text}  [{ops}]")
    out = "\n".join(lines) + "\n"
    if len(sys.argv) > 4:
        open(sys.argv[4], "w").write(out)
    else:
        sys.stdout.write(out)


if __name__ == "__main__":
    main()
