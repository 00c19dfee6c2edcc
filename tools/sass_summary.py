"""SASS evidence for profiles/sass/: per-kernel mnemonic counts of the in-tree libb200codec.so and full listings of
the hot kernels (run here, no GPU needed):

    python tools/sass_summary.py --tag r02

tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UBLKCP, mma.sync -> HMMA
(/opt/skills/guides/B200_PROFILING.md "What proves a Blackwell-native kernel").
"""
import argparse
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tts_max_b200", "lib", "libb200codec.so")
COLS = ["UTCHMMA", "UTCBAR", "UTMALDG", "UTMASTG", "LDTM", "STTM", "HMMA", "MUFU", "FFMA2/FADD2/FMUL2", "ATOM/RED", "BAR/SYNCS"]
FULL = {"gemm_tc05_2cta_kernel<__nv_bfloat16, false, 256, false>": "gemm_tc05_2cta_bf16_n256",
        "attention_tc05_kernel<__nv_bfloat16>": "attention_tc05_bf16",
        "istft_kernel<320, 16>": "istft_hop320_16warps",
        "snake_aa_kernel<__nv_bfloat16, 12>": "snake_aa_bf16_67rows",
        "gemm_tc05_2cta_kernel<__nv_bfloat16, false, 192, false>": "gemm_tc05_2cta_bf16_n192"}

ap = argparse.ArgumentParser()
ap.add_argument("--tag", default="r02")
args = ap.parse_args()
out_dir = os.path.join(ROOT, "profiles", "sass")
os.makedirs(out_dir, exist_ok=True)
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
funcs = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
    elif cur is not None:
        funcs[cur].append(line)
names = subprocess.run(["c++filt"], input="\n".join(funcs), capture_output=True, text=True).stdout.splitlines()


def short(n):
    n = re.sub(r"void |b200::|\(anonymous namespace\)::", "", n)
    n = re.sub(r"\(.*", "", n)
    return n.replace("(bool)0", "false").replace("(bool)1", "true").replace("(int)", "")


rows = []
for (mangled, body), dem in zip(funcs.items(), names):
    ops = [m.group(1) for l in body for m in [re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", l)] if m]
    base = collections.Counter(o.split(".")[0] for o in ops)
    cnt = {"UTCHMMA": base["UTCHMMA"], "UTCBAR": base["UTCBAR"], "UTMALDG": base["UTMALDG"], "UTMASTG": base["UTMASTG"],
           "LDTM": base["LDTM"], "STTM": base["STTM"], "HMMA": base["HMMA"], "MUFU": base["MUFU"],
           "FFMA2/FADD2/FMUL2": base["FFMA2"] + base["FADD2"] + base["FMUL2"],
           "ATOM/RED": base["ATOM"] + base["ATOMG"] + base["RED"] + base["REDG"] + base["ATOMS"],
           "BAR/SYNCS": base["BAR"] + base["SYNCS"]}
    name = short(dem)
    rows.append((name, len(ops), cnt))
    if name in FULL:
        path = os.path.join(out_dir, f"{args.tag}_{FULL[name]}.sass")
        with open(path, "w") as f:
            f.write(f"// {dem}\n// cuobjdump -sass tts_max_b200/lib/libb200codec.so (sm_100a), {len(ops)} instructions\n")
            # drop the hex encodings (second comment of every line, and the encoding-only lines): half the size
            keep = [re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l) for l in body]
            f.write("\n".join(l for l in keep if l.strip()) + "\n")
git = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
lines = [f"# {args.tag}: SASS mnemonic counts per kernel of tts_max_b200/lib/libb200codec.so (sm_100a), sources at {git}", "",
         "`python tools/sass_summary.py --tag " + args.tag + "` (cuobjdump -sass | c++filt). tcgen05.mma = UTCHMMA, tcgen05.commit = UTCBAR, "
         "TMA load / store = UTMALDG / UTMASTG, tcgen05.ld / st = LDTM / STTM; an HMMA (mma.sync) count other than 0 would be a "
         "legacy tensor-core path. Full listings: " + ", ".join(f"`profiles/sass/{args.tag}_{v}.sass`" for v in FULL.values()) + ".", "",
         "| kernel | instr | " + " | ".join(COLS) + " |", "|---|---:|" + "---:|" * len(COLS)]
for name, n, cnt in sorted(rows, key=lambda r: r[0]):
    lines.append(f"| `{name}` | {n} | " + " | ".join(str(cnt[c]) for c in COLS) + " |")
with open(os.path.join(out_dir, f"{args.tag}_sass_summary.md"), "w") as f:
    f.write("\n".join(lines) + "\n")
print("\n".join(lines[:6]))
print(len(rows), "kernels")
