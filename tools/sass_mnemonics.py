"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native path (tcgen05 -> UTC*MMA, tcgen05.ld/st ->
LDTM/STTM, TMA -> UTMALDG/UTMASTG/UTMAREDG, commits -> UTCBAR) and of the legacy tensor-core ones that must not
appear (HMMA., WGMMA/HGMMA).  Usage: python tools/sass_mnemonics.py > profiles/r2_sass_mnemonics.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "clip_lora_match_b200", "csrc", "libclm_b200.so")
WANT = ["UTCHMMA.2CTA", "UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "UTMAPF", "MUFU.EX2",
        "MUFU.TANH", "FMNMX3", "HMMA.", "HGMMA", "WGMMA"]

def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, name = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(CUtensorMap_st.*", "", name).replace("(anonymous namespace)::", "")
            counts[name] = collections.Counter()
            continue
        if name is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        for w in WANT:
            if op.startswith(w):
                if w == "UTCHMMA" and op.startswith("UTCHMMA.2CTA"):
                    continue
                counts[name][w] += 1
                break
    print(f"# SASS mnemonic counts per kernel of {os.path.relpath(LIB, ROOT)} (cuobjdump -sass); zero columns omitted")
    tot = collections.Counter()
    for k, c in counts.items():
        if not c:
            continue
        tot.update(c)
        print(f"{k}\n    " + "  ".join(f"{w}={c[w]}" for w in WANT if c[w]))
    print("TOTAL\n    " + "  ".join(f"{w}={tot[w]}" for w in WANT))
    legacy = tot["HMMA."] + tot["HGMMA"] + tot["WGMMA"]
    print(f"legacy tensor-core instructions (HMMA. / HGMMA / WGMMA): {legacy}")
    sys.exit(1 if legacy else 0)

if __name__ == "__main__":
    main()
