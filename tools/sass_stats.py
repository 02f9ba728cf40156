"""Static SASS statistics of a generated kernel: instructions in the step loop (the largest
backward-branch body) and an opcode histogram.  The rollout kernels are issue-bound and their
steps are (almost) branch-free, so this count tracks executed instructions per warp-step."""
import collections
import re
import subprocess
import sys


def sass(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    ins = []
    for line in out.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    return ins


def loop_body(ins):
    """the largest [target, branch] span of a backward branch"""
    best = None
    for i, (addr, text) in enumerate(ins):
        m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", text)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < addr:
                j = next(k for k, (a, _) in enumerate(ins) if a == tgt)
                if best is None or i - j > best[1] - best[0]:
                    best = (j, i)
    return best


def main(path, top=25):
    ins = sass(path)
    lb = loop_body(ins)
    print(f"{path}: {len(ins)} SASS instructions; step loop {lb[1] - lb[0] + 1 if lb else 0}")
    if lb:
        body = ins[lb[0]:lb[1] + 1]
        hist = collections.Counter()
        for _, t in body:
            t = re.sub(r"^@!?U?P\d+\s+", "", t)
            hist[t.split()[0].split(".")[0]] += 1
        print("  " + ", ".join(f"{k} {v}" for k, v in hist.most_common(top)))
        nbr = sum(1 for _, t in body if re.search(r"\bBRA\b|\bBSSY\b|\bCALL\b", t))
        print(f"  branches/calls in loop: {nbr}")


if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
