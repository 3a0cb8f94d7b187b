"""Print the top stalled SASS lines of each kernel in an `ncu --page source --csv` dump."""
import csv
import sys


def main(path, which=0, n=40):
    r = list(csv.reader(open(path)))
    blocks, cur = [], None
    for x in r:
        if x and x[0] == "Kernel Name":
            cur = dict(name=x[1], hdr=None, rows=[])
            blocks.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = x
        elif cur is not None and len(x) == len(cur["hdr"]):
            cur["rows"].append(x)
    print(len(blocks), "kernels")
    b = blocks[which]
    hdr, rows = b["hdr"], b["rows"]
    si, src, ie = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(x[si]) for x in rows)
    print(b["name"][:100], "total samples", tot, "instructions", len(rows))
    top = sorted(range(len(rows)), key=lambda i: -int(rows[i][si]))[:n]
    for i in sorted(top):
        x = rows[i]
        st = {hdr[c][6:]: int(x[c]) for c in stall_cols if int(x[c]) > 0}
        st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
        print(i, x[si], x[ie], x[src].strip()[:80], st)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else 40)
