"""Hot spots of a kernel from an exported ncu source page (SASS view, `ncu -i rep --page source --csv`, gzipped): consecutive SASS
instructions with the same execution count are grouped into blocks; prints the blocks by share of executed warp instructions with
their opcode mix, average active threads and stall samples.   python scripts/ncu_sass_blocks.py profiles/<file>.csv.gz [top=28]"""
import collections, csv, gzip, io, sys


def num(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


def main(path, top=28):
    txt = gzip.open(path, "rt").read().splitlines()
    i = next(k for k, l in enumerate(txt) if l.startswith('"') and "Source" in l)
    rows = list(csv.reader(io.StringIO("\n".join(txt[i:]))))
    cols = rows[0]
    rows = [r for r in rows[1:] if len(r) == len(cols) and r[0].startswith("0x")]
    ie, ta, ss = cols.index("Instructions Executed"), cols.index("Avg. Threads Executed"), cols.index("# Samples")

    def opcode(src):
        t = src.split()
        o = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
        return o.split(".")[0]
    blocks, cur = [], None
    for k, r in enumerate(rows):
        c = num(r[ie])
        if cur and abs(c - cur["c"]) <= 0.02 * max(c, cur["c"], 1):
            cur["n"] += 1; cur["tot"] += c; cur["samples"] += num(r[ss]); cur["ops"].append(opcode(r[1])); cur["thr"] += num(r[ta])
        else:
            cur = {"start": k, "c": c, "n": 1, "tot": c, "samples": num(r[ss]), "ops": [opcode(r[1])], "thr": num(r[ta])}
            blocks.append(cur)
    tot = sum(b["tot"] for b in blocks)
    tots = sum(b["samples"] for b in blocks)
    print(f"{path}: {len(rows)} SASS instructions, {tot:.3e} executed warp instructions, {tots:.0f} stall samples")
    for b in sorted(blocks, key=lambda b: -b["tot"])[:top]:
        oc = collections.Counter(b["ops"])
        print(f"  sass[{b['start']:5d}..{b['start'] + b['n'] - 1:5d}] exec/instr {b['c']:12.0f}  share {100 * b['tot'] / tot:5.1f}%  threads {b['thr'] / b['n']:4.1f}  "
              f"samples {100 * b['samples'] / max(tots, 1):5.1f}%  {dict(oc.most_common(8))}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 28)
