"""Top stall sites of one kernel from an `ncu --page source --csv` dump: samples per SASS instruction, with neighbours."""
import csv
import sys


def main(path, kernel_regex_hint="", topn=25):
    with open(path, newline="") as f:
        rows = list(csv.reader(f))
    # find header rows (one per kernel instance); use the first instance only
    start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[start]
    end = next((i for i in range(start + 1, len(rows)) if rows[i] and rows[i][0] in ("Kernel Name", "Address")), len(rows))
    body = [r for r in rows[start + 1:end] if len(r) >= len(hdr) - 5]
    si, ai, ni = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("# Samples")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") or h.startswith("Stall")]
    total = sum(int(r[ai] or 0) for r in body)
    print(f"{len(body)} SASS instructions, {total} stall samples")
    order = sorted(range(len(body)), key=lambda i: -int(body[i][ai] or 0))[:topn]
    for i in order:
        r = body[i]
        reasons = sorted(((int(r[c] or 0), hdr[c]) for c in stall_cols if (r[c] or "0").isdigit() and int(r[c] or 0) > 0), reverse=True)[:3]
        print(f"{int(r[ai]) / max(total, 1) * 100:5.1f}%  #{i:5d}  {r[si].strip()[:90]:90s} {reasons}")


if __name__ == "__main__":
    main(sys.argv[1], topn=int(sys.argv[2]) if len(sys.argv) > 2 else 25)
