"""Summarise `nvcc -Xptxas -v` output: kernel, registers, spills, smem (reads stdin or a file)."""
import re
import subprocess
import sys


def demangle(names):
    try:
        out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
        return out[:len(names)]
    except Exception:
        return names


def main():
    text = open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read()
    rows = []
    cur = None
    for line in text.splitlines():
        m = re.search(r"Compiling entry function '([^']+)'", line)
        if m:
            cur = dict(name=m.group(1), regs=None, spill_st=0, spill_ld=0, smem=0)
            rows.append(cur)
            continue
        if cur is None:
            continue
        m = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m:
            cur["spill_st"], cur["spill_ld"] = int(m.group(1)), int(m.group(2))
        m = re.search(r"Used (\d+) registers", line)
        if m:
            cur["regs"] = int(m.group(1))
            m2 = re.search(r"(\d+) bytes smem", line)
            if m2:
                cur["smem"] = int(m2.group(1))
    names = demangle([r["name"] for r in rows])
    for r, n in zip(rows, names):
        n = re.sub(r"\(anonymous namespace\)::", "", n)
        n = re.sub(r"\(.*", "", n)
        if len(sys.argv) > 2 and not re.search(sys.argv[2], n):
            continue
        print("%-70s regs=%-4s spill=%d/%d smem=%d" % (n[:70], r["regs"], r["spill_st"], r["spill_ld"], r["smem"]))


if __name__ == "__main__":
    main()
