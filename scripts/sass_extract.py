"""profiles/r2_sass_extract.md from the shipped library: `cuobjdump -sass lora_b200/liblora_b200.so`,
per-kernel counts of the Blackwell-native mnemonics and the first occurrences with context.
Runs on the build machine (no GPU needed):  python scripts/sass_extract.py > profiles/r2_sass_extract.md"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "lora_b200", "liblora_b200.so")
PATS = [("UTCHMMA", r"\bUTCHMMA"), ("LDTM", r"\bLDTM"), ("UTCBAR", r"\bUTCBAR"), ("UTMALDG", r"\bUTMALDG"),
        ("UTMASTG", r"\bUTMASTG"), ("UBLKCP (DSMEM bulk copy)", r"\bUBLKCP"), ("REDG.F32x4", r"\bREDG\.E\.ADD\.F32x4|RED\.E\.ADD\.F32x4"),
        ("HMMA", r"\bHMMA"), ("LDSM", r"\bLDSM"), ("LDGSTS", r"\bLDGSTS")]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    txt = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    funcs, cur, name = {}, None, None
    for ln in txt.split("\n"):
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            name = m.group(1)
            cur = funcs.setdefault(name, [])
            continue
        if cur is not None and "/*" in ln and ";" in ln:
            cur.append(re.sub(r"/\*[0-9a-fx]+\*/", "", ln).strip())
    dm = demangle(list(funcs))
    short = lambda n: re.sub(r"\(.*", "", dm[n])[:110]
    print("# SASS evidence (round 2, final build): `cuobjdump -sass lora_b200/liblora_b200.so` (scripts/sass_extract.py)\n")
    print("Per kernel: instruction counts of the Blackwell-native mnemonics (tcgen05.mma = `UTCHMMA`, tcgen05.ld = `LDTM`,")
    print("tcgen05.commit = `UTCBAR`, TMA loads/stores = `UTMALDG` / `UTMASTG`, the cluster split-K kernel's distributed-shared-memory")
    print("bulk copy = `UBLKCP`, vector fp32 reductions of the L2 split-K path = `REDG.E.ADD.F32x4`, legacy tensor path of the SVD")
    print("passes = `HMMA`, their cp.async ring = `LDGSTS`), then first occurrences with two lines of context.\n")
    print("| kernel | " + " | ".join(p[0] for p in PATS) + " |")
    print("|---|" + "---|" * len(PATS))
    rows = []
    for n, body in funcs.items():
        cnt = [sum(1 for l in body if re.search(p[1], l)) for p in PATS]
        if sum(cnt[:7]) + cnt[7] + cnt[9] == 0:
            continue
        rows.append((short(n), cnt, n))
    rows.sort()
    for s, cnt, _ in rows:
        print(f"| `{s}` | " + " | ".join(str(c) for c in cnt) + " |")
    print()
    want = ["fused_lora_kernel<64, 7, unsigned short, false, 1, 1, false, false", "fused_lora_splitk_kernel<64, 6, unsigned short, 2",
            "fused_lora_kernel<192, 4, unsigned short", "fused_lora_persistent_kernel<128, 4, unsigned short",
            "fused_lora_kernel<128, 4, unsigned short, false, 1, 1, false, true", "mul_right_kernel<__half, 2"]
    for w in want:
        for s, cnt, n in rows:
            if w in s:
                body = funcs[n]
                print(f"## `{s}`  ({len(body)} instructions)\n\n```")
                shown = set()
                for pname, pat in PATS:
                    for i, l in enumerate(body):
                        if re.search(pat, l) and pname not in shown:
                            shown.add(pname)
                            for j in range(max(0, i - 2), min(len(body), i + 3)):
                                print(("  > " if j == i else "    ") + body[j])
                            print("    ...")
                            break
                print("```\n")
                break


if __name__ == "__main__":
    main()
