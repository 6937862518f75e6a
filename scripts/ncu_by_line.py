"""Joins an ncu SASS source page (ncu -i X.ncu-rep --page source --csv --print-source sass) with nvdisasm -g line info of the same
kernel and aggregates executed instructions / stall samples by source function and by source line.
  python scripts/ncu_by_line.py <sass.csv> <nvdisasm -g -c output> <mangled kernel name> [top]"""
import collections
import csv
import re
import sys


def load_lines(dis_path, kernel):
    out = []
    active = False
    cur = ("?", 0)
    for line in open(dis_path, errors="replace"):
        if line.startswith(".text."):
            active = line.strip() == ".text.%s:" % kernel
            continue
        if not active:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
        if m:
            if "inlined at" not in line:
                cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            out.append((int(m.group(1), 16), cur, m.group(2).strip()))
    return out


def function_of(cache, path_by_name, fname, line):
    if fname not in cache:
        starts = []
        try:
            src = open(path_by_name.get(fname, fname), errors="replace").read().split("\n")
        except OSError:
            src = []
        for i, l in enumerate(src, 1):
            m = re.match(r"^\s*(?:template\s*<[^>]*>\s*)?(?:UPC_\w+|inline|static|__global__|__device__|__host__|__forceinline__|\w+)[\w\s\*&:<>,]*?\b(\w+)\s*\([^;]*$", l)
            if m and not l.strip().startswith(("if", "for", "while", "return", "else", "switch", "//", "*", "#")) and "=" not in l.split("(")[0]:
                starts.append((i, m.group(1)))
        cache[fname] = starts
    name = "?"
    for i, n in cache[fname]:
        if i <= line:
            name = n
        else:
            break
    return name


def main():
    sass_csv, dis, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
    lines = load_lines(dis, kernel)
    rows = list(csv.reader(open(sass_csv)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
    h = rows[hdr]
    ci, si = h.index("Instructions Executed"), h.index("# Samples")
    src_i = h.index("Source")
    body = rows[hdr + 1:]
    n = min(len(body), len(lines))
    import glob
    import os
    # UPC_SRC_ROOT: the checkout the profiled binary was built from (line numbers must match), default = this tree
    root = os.environ.get("UPC_SRC_ROOT") or os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path_by_name = {os.path.basename(p): p for p in glob.glob(root + "/uncertainty_planning_core_b200/csrc/*") + glob.glob(root + "/include/*")}
    by_fn, by_line = collections.Counter(), collections.Counter()
    local_fn = collections.Counter()   # executed local-memory loads / stores (LDL / STL) per function
    tot_local = 0
    s_fn, s_line = collections.Counter(), collections.Counter()
    cache = {}
    tot_i = tot_s = 0
    for k in range(n):
        try:
            cnt = int(float(body[k][ci] or 0))
            smp = int(float(body[k][si] or 0))
        except ValueError:
            continue
        f, l = lines[k][1]
        fn = "%s:%s" % (f, function_of(cache, path_by_name, f, l))
        by_fn[fn] += cnt
        s_fn[fn] += smp
        op = body[k][src_i].strip().split()
        op = op[1] if op and op[0].startswith("@") and len(op) > 1 else (op[0] if op else "")
        if op.startswith("LDL") or op.startswith("STL"):
            local_fn[fn] += cnt
            tot_local += cnt
        by_line[(f, l)] += cnt
        s_line[(f, l)] += smp
        tot_i += cnt
        tot_s += smp
    print("instructions %d, samples %d, sass rows %d / disasm rows %d" % (tot_i, tot_s, len(body), len(lines)))
    print("local-memory instructions (LDL + STL) %d = %.1f %% of all" % (tot_local, 100.0 * tot_local / max(tot_i, 1)))
    print("%-55s %8s %8s %9s" % ("function", "instr %", "stall %", "LDL+STL %"))
    for fn, c in by_fn.most_common(top):
        print("%-55s %7.1f%% %7.1f%% %8.1f%%" % (fn, 100.0 * c / max(tot_i, 1), 100.0 * s_fn[fn] / max(tot_s, 1), 100.0 * local_fn[fn] / max(tot_local, 1)))
    print("\n%-40s %8s %8s" % ("line", "instr %", "stall %"))
    for (f, l), c in by_line.most_common(top):
        print("%-40s %7.1f%% %7.1f%%" % ("%s:%d" % (f, l), 100.0 * c / max(tot_i, 1), 100.0 * s_line[(f, l)] / max(tot_s, 1)))


if __name__ == "__main__":
    main()
