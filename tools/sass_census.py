"""Instruction census of the tcgen05 kernels of libdnmf.so (cuobjdump -sass): proof that the hot kernels are tcgen05 / TMA /
TMEM code.    python tools/sass_census.py > profiles/r02_sass_census.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ('UTCHMMA', 'UTMALDG', 'LDTM', 'STTM', 'UTCBAR', 'FADD2', 'FMUL2', 'MUFU.RCP', 'SYNCS.PHASECHK', 'NANOSLEEP')


def main():
    out = subprocess.run(['cuobjdump', '-sass', os.path.join(ROOT, 'pydnmfk_b200', 'libdnmf.so')], capture_output=True, text=True).stdout
    print('# SASS instruction census of the tcgen05 kernels in pydnmfk_b200/libdnmf.so (cuobjdump -sass, sm_100a), final build of round 2')
    print('# UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, FADD2 / FMUL2 = packed fp32')
    print('# (two tc_kl_kernel<0|1> entries: the 32-wide build of dnmf_tc_kl.cu and the 64-wide build of dnmf_tc_kl64.cu)')
    name, n, cnt = None, 0, collections.Counter()

    def flush():
        if name and ('tc_pass_kernel' in name or 'tc_kl_kernel' in name):
            short = re.search(r'(tc_(?:pass|kl)_kernelI[A-Za-z0-9]*?E)(?:EEv|Ev)', name)
            print('%-30s %5d instructions  %s' % (short.group(1) if short else name[:30], n,
                                                  ' '.join('%s=%d' % (k, cnt[k]) for k in sorted(cnt))))
    for line in out.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            flush()
            name, n, cnt = m.group(1), 0, collections.Counter()
            continue
        m = re.match(r'\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if m:
            n += 1
            for k in KEYS:
                if m.group(1).startswith(k):
                    cnt[k] += 1
    flush()


if __name__ == '__main__':
    main()
