"""SASS evidence of the shipped library: counts of the Blackwell-native mnemonics per kernel + an excerpt.
Runs without a GPU:  python tools/sass_summary.py > profiles/r02_sass_summary.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'patch2pix_b200', 'libp2p_b200.so')
PAT = {'tcgen05.mma (UTC*MMA)': r'\bUTC[A-Z]*MMA', 'tcgen05.mma cta_group::2 (.2CTA)': r'UTC[A-Z]*MMA\.2CTA', 'TMA load (UTMALDG)': r'\bUTMALDG', 'TMA store (UTMASTG)': r'\bUTMASTG', 'bulk copy (UBLKCP)': r'\bUBLKCP',
       'tcgen05.ld (LDTM)': r'\bLDTM', 'tcgen05.commit (UTCBAR)': r'\bUTCBAR', 'TMEM alloc (UTCATOMSWS)': r'\bUTCATOMSWS',
       'mbarrier (SYNCS)': r'\bSYNCS', 'legacy HMMA (mma.sync)': r'\bHMMA', 'FFMA': r'\bFFMA'}


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    excerpt = {}
    for ln in sass.splitlines():
        m = re.search(r'Function : (\S+)', ln)
        if m:
            cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip().split('(')[0]
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for name, pat in PAT.items():
            if re.search(pat, ln):
                per[cur][name] += 1
                if name.startswith(('tcgen05.mma (', 'TMA load', 'tcgen05.ld')) and len(excerpt.setdefault(cur, [])) < 6:
                    excerpt[cur].append(ln.strip()[:150])
    print('# SASS summary of `patch2pix_b200/libp2p_b200.so` (cuobjdump -sass, sm_100a)\n')
    print('| kernel | ' + ' | '.join(PAT) + ' |')
    print('|---|' + '---|' * len(PAT))
    tot = collections.Counter()
    for k, c in per.items():
        if not any(c[n] for n in PAT if n != 'FFMA'):
            continue
        tot.update(c)
        print(f'| `{k[:90]}` | ' + ' | '.join(str(c[n]) for n in PAT) + ' |')
    print('| **total** | ' + ' | '.join(str(tot[n]) for n in PAT) + ' |')
    print('\n## Excerpts (first tensor-core / TMA / TMEM instructions per kernel)\n')
    for k, lines in excerpt.items():
        print(f'### `{k[:110]}`\n```')
        print('\n'.join(lines))
        print('```')


if __name__ == '__main__':
    main()
