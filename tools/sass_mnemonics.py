"""cuobjdump -sass of the built library -> counts of the tensor / TMA / TMEM mnemonics per kernel family (profiles/*_sass_mnemonics.txt):
    python tools/sass_mnemonics.py > profiles/rNN_sass_mnemonics.txt"""
import subprocess, re, collections
out = subprocess.run(["cuobjdump", "-sass", "mraudio_b200/libmraudio_b200.so"], capture_output=True, text=True).stdout
fam = None
per = collections.defaultdict(collections.Counter)
tot = collections.Counter()
pat = re.compile(r"\b(UTCHMMA[.\w]*|UTMALDG[.\w]*|UTMASTG[.\w]*|UTMAREDG[.\w]*|LDTM[.\w]*|UTCBAR[.\w]*|HMMA[.\w]*|LDSM[.\w]*|UTCATOMSWS[.\w]*|MUFU\.TANH)")
for line in out.splitlines():
    if "Function :" in line:
        m = re.search(r"(gemm_tc_kernel|gemm_ln_kernel|qkv_attn_kernel|attention_tma_kernel|attention_kernel|attn_bwd_tc_kernel)", line)
        fam = m.group(1) if m else "other"
        continue
    m = pat.search(line)
    if m and fam:
        per[fam][m.group(1)] += 1
        tot[m.group(1)] += 1
print("# cuobjdump -sass mraudio_b200/libmraudio_b200.so (final round-2 build): tensor / TMA / TMEM mnemonics")
print("## whole library")
for k, v in tot.most_common():
    print(f"{v:7d} {k}")
for f in ("qkv_attn_kernel", "gemm_tc_kernel", "gemm_ln_kernel", "attention_tma_kernel", "attn_bwd_tc_kernel"):
    print(f"## {f} (all template instances)")
    for k, v in per[f].most_common():
        print(f"{v:7d} {k}")
