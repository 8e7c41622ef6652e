"""Regenerate include/airdos_orb_pattern.h from the reference's embedded rBRIEF table.

Runs only where /root/reference is mounted (the build container); the generated header is
committed.  The table is the learned ORB sampling pattern (src/ORBextractor.cc:151-409).
"""
import re, sys

def main(ref='/root/reference/src/ORBextractor.cc', out='include/airdos_orb_pattern.h'):
    src = open(ref).read()
    i = src.index('bit_pattern_31_[256*4]'); j = src.index('};', i)
    body = re.sub(r'/\*.*?\*/', '', src[src.index('{', i) + 1:j], flags=re.S)
    nums = [int(t) for t in re.findall(r'-?\d+', body)]
    assert len(nums) == 1024
    xs, ys = nums[0::2], nums[1::2]
    def fmt(v):
        return ' \\\n'.join('    ' + ','.join('%d' % t for t in v[k:k + 32]) + ',' for k in range(0, 512, 32)) + ' \\'
    hdr = open(out).read().split('#define AIRDOS_ORB_PATTERN_X')[0]
    open(out, 'w').write(hdr + '#define AIRDOS_ORB_PATTERN_X { \\\n%s\n }\n#define AIRDOS_ORB_PATTERN_Y { \\\n%s\n }\n#endif\n' % (fmt(xs), fmt(ys)))

if __name__ == '__main__':
    main(*sys.argv[1:])
