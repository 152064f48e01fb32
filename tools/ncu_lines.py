"""Per-source-line instruction / stall-sample shares of one kernel in an .ncu-rep.  Usage: ncu_lines.py rep kernel-regex [top]"""
import csv, io, subprocess, sys


def src(rep, kern, top=25):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv', '-k', 'regex:' + kern],
                         capture_output=True, text=True).stdout
    cur, agg = None, []
    for r in csv.reader(io.StringIO(out)):
        if len(r) >= 2 and r[0] == 'File Path':
            cur = r[1].split('/')[-1]
        elif len(r) >= 8 and r[0].isdigit():
            try:
                agg.append((int(r[7]), int(r[4]) if r[4].isdigit() else 0, cur, int(r[0]), r[1].strip()[:110]))
            except ValueError:
                pass
    ti, ts = sum(a[0] for a in agg) or 1, sum(a[1] for a in agg) or 1
    for a in sorted(agg, key=lambda a: -a[1])[:top]:
        print(f"{a[0] / ti * 100:5.1f}% inst {a[1] / ts * 100:5.1f}% samples  {a[2]}:{a[3]}  {a[4]}")


if __name__ == '__main__':
    src(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25)
