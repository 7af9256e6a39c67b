"""profiles/sass_summary.txt: SASS instruction counts of the built library (the tcgen05 / TMA / mbarrier mnemonics of
/opt/skills/guides/B200_PROFILING.md), per kernel for the tensor-core instructions.  python scripts/sass_summary.py > profiles/sass_summary.txt"""
import collections
import hashlib
import os
import re
import subprocess

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, 'ood_gan_inversion_b200', 'libood_b200.so')
sass = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True, check=True).stdout
digest = hashlib.sha256(open(so, 'rb').read()).hexdigest()[:16]
kernels = sass.count('Function :')
pats = [('UTCHMMA (tcgen05.mma), all', r'\bUTCHMMA\b'), ('UTCHMMA.2CTA (tcgen05.mma.cta_group::2)', r'\bUTCHMMA\.2CTA'),
        ('UTCBAR (tcgen05.commit), all', r'\bUTCBAR\b'), ('UTCBAR.2CTA.MULTICAST (commit to both CTAs of a pair)', r'UTCBAR\.2CTA\.MULTICAST'),
        ('LDTM (tcgen05.ld)', r'\bLDTM'), ('UTMALDG (TMA load), all', r'\bUTMALDG'), ('UTMALDG.*.2CTA (TMA load, completion on the pair leader)', r'UTMALDG\.\dD\.2CTA'),
        ('UTMASTG (TMA store)', r'\bUTMASTG'), ('UTMAPF / UBLKPF (TMA prefetch)', r'\bUTMAPF|\bUBLKPF'), ('SYNCS (mbarrier)', r'\bSYNCS'),
        ('ELECT', r'\bELECT\b'), ('R2UR', r'\bR2UR\b'), ('HMMA (mma.sync, ToRGB only)', r'\bHMMA'),
        ('FFMA2/FMUL2/FADD2 (packed fp32x2)', r'\bFFMA2|\bFMUL2|\bFADD2'), ('LDGSTS (cp.async)', r'\bLDGSTS'), ('STG.E.ENL2.256 / STG.256', r'STG\.[A-Z0-9.]*256'),
        ('UCGABAR (cluster barrier)', r'\bUCGABAR')]
print('SASS instruction counts of ood_gan_inversion_b200/libood_b200.so (cuobjdump -sass), round 2')
print(f'built by ood_gan_inversion_b200/build.py: nvcc 12.9 -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo; {kernels} kernels; sha256[:16] of this build {digest}')
print('reproduce: python -c "import __graft_entry__ as g; g.build()" && python scripts/sass_summary.py\n')
for name, pat in pats:
    print(f'{len(re.findall(pat, sass)):7d}  {name}')
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        cur = m.group(1)
        continue
    if cur and re.search(r'\bUTCHMMA', line):
        d = per.setdefault(cur, [0, 0])
        d[0] += 1
        d[1] += 1 if 'UTCHMMA.2CTA' in line else 0
names = subprocess.run(['c++filt'], input='\n'.join(per), capture_output=True, text=True).stdout.splitlines()
print('\nkernels that issue tcgen05.mma (UTCHMMA count per kernel, of which .2CTA):')
for (k, (n, n2)), name in zip(per.items(), names):
    print(f'{n:5d} {n2:4d}  {name}')
