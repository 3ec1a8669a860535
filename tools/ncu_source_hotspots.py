"""Source-page CSV of one kernel of an `ncu --set full --import-source on` capture -> where the warps of that kernel wait.

    ncu -i gpurun_out/r02_full_65536.ncu-rep --page source --csv --print-source sass,cuda \
        --kernel-id ::regex:tc_kl_kernel:1 > /tmp/kl0_source.csv
    python tools/ncu_source_hotspots.py /tmp/kl0_source.csv [top]

Prints (1) the stall-reason totals of the kernel (not-issued warp samples), (2) the CUDA source lines with the most samples,
each with its dominant reasons and the SASS opcodes that collected them, (3) the instruction mix by opcode class.  Samples
are per warp: a kernel with role-specialised warps shows each role's wait where that role's code is."""
import collections
import csv
import re
import sys


def _int(x):
    try:
        return int(float(x.replace(',', '')))
    except ValueError:
        return 0


def main(path, top=22):
    rows = list(csv.reader(open(path)))
    print(next((r[1][:110] for r in rows if r and r[0] == 'Function Name'), ''))
    lines, cur, cur_file, hdr = collections.OrderedDict(), None, '', None
    totals, opmix, n_inst, seen, insts, cur_line = collections.Counter(), collections.Counter(), 0, set(), [], 0
    for r in rows:                                                   # one section per source file the kernel inlines from
        if not r:
            continue
        if r[0] == 'File Path':
            cur_file, cur = r[1].rsplit('/', 1)[-1], None
            continue
        if r[0] == 'Function Name':
            continue
        if r[0] == 'Line No':
            hdr = r
            col = {name: i for i, name in enumerate(hdr)}
            i_samples, i_exec = col['# Samples'], col['Instructions Executed']
            stall_cols = [(n[:-len(' (Not Issued)')], i) for n, i in col.items()
                          if n.startswith('stall_') and n.endswith('(Not Issued)')]
            src_i = [i for i, n in enumerate(hdr) if n == 'Source']  # first = CUDA line text, second = SASS text
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        if r[0]:                                                     # a CUDA source line (aggregated) ...
            cur_line = int(r[0])
            cur = lines.setdefault((cur_file, int(r[0])), {'text': r[src_i[0]].strip(), 'samples': 0, 'exec': 0,
                                                           'stalls': collections.Counter(), 'ops': collections.Counter()})
            continue
        if cur is None:                                              # ... followed by its SASS instructions
            continue
        addr = r[col['Address']]
        if not addr.startswith('0x') or addr in seen:                                             # the view lists every instruction of a line twice
            continue
        seen.add(addr)
        sass = r[src_i[1]].strip()
        insts.append((int(addr, 16), sass, _int(r[i_samples]), cur_file, cur_line))
        op = re.sub(r'^@!?U?P\d+\s+', '', sass).split(' ')[0].split('.')[0]
        s = _int(r[i_samples])
        e = _int(r[i_exec])
        cur['samples'] += s
        cur['exec'] += e
        cur['ops'][op] += s
        opmix[op] += e
        n_inst += e
        for name, i in stall_cols:
            v = _int(r[i])
            if v:
                cur['stalls'][name] += v
                totals[name] += v
    all_samples = sum(v['samples'] for v in lines.values())
    not_issued = sum(totals.values())
    print('warp samples: %d (not issued: %d); warp-instructions executed: %.3g' % (all_samples, not_issued, n_inst))
    print('\nstall reasons (share of not-issued samples):')
    for name, v in totals.most_common(10):
        print('  %-26s %5.1f %%' % (name, 100.0 * v / max(not_issued, 1)))
    print('\nsource lines by samples:')
    print('  %-22s %6s %9s  %-44s %s' % ('file:line', 'share', 'inst', 'dominant stalls', 'source'))
    for ln, v in sorted(lines.items(), key=lambda kv: -kv[1]['samples'])[:top]:
        st = ', '.join('%s %d%%' % (n.replace('stall_', ''), round(100.0 * c / max(sum(v['stalls'].values()), 1)))
                       for n, c in v['stalls'].most_common(3))
        ops = '/'.join(o for o, _ in v['ops'].most_common(3))
        print('  %-22s %5.1f%% %9.3g  %-44s %s   [%s]' % ('%s:%d' % ln, 100.0 * v['samples'] / max(all_samples, 1), v['exec'], st,
                                                      v['text'][:70], ops))
    # the hottest single instructions, each with the range of main-file lines of the code around it (= which role's loop)
    insts.sort()
    main_file = collections.Counter(f for _, _, _, f, _ in insts).most_common(1)[0][0]
    order = sorted(range(len(insts)), key=lambda i: -insts[i][2])[:14]
    print('\nhottest instructions (samples, context = main-file lines within +-60 instructions):')
    for i in order:
        ctx = [ln for _, _, _, f, ln in insts[max(0, i - 60):i + 60] if f == main_file and ln > 200]
        print('  %5.1f%%  %-46s %s:%d  context %s:%s-%s' % (100.0 * insts[i][2] / max(all_samples, 1), insts[i][1][:46], insts[i][3],
                                                          insts[i][4], main_file, min(ctx) if ctx else '?', max(ctx) if ctx else '?'))
    print('\nwarp-instruction mix:')
    for op, e in opmix.most_common(16):
        print('  %-10s %5.1f %%' % (op, 100.0 * e / max(n_inst, 1)))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 22)
