"""Offline SASS view of one kernel: `python tools/sass_loop.py obj.o <function substring> [first-line last-line]` prints the
instruction stream without encodings (for counting the instructions of a loop body before spending GPU time)."""
import subprocess, sys, re
obj, fn = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout.splitlines()
on, rows = False, []
for l in out:
    if "Function :" in l:
        on = fn in l
        continue
    if on:
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m:
            rows.append((m.group(1), m.group(2).strip()))
lo = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4]) if len(sys.argv) > 4 else len(rows)
for i, (a, t) in enumerate(rows[lo:hi], lo):
    print(i, a, t)
print("total", len(rows), file=sys.stderr)
