"""Per-layer efficiency of the trunk convs from an ncu launch list (largest-scale forward). Development aid."""
import sys
sys.path.insert(0, 'tools')
from launch_summary import load
L = load(sys.argv[1])
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
h0 = int(sys.argv[3]) if len(sys.argv) > 3 else 256
# locate the forward with the largest first-layer conv: find max duration conv_tc<64> with grid (B*h0*h0/128)
tiles = B * h0 * h0 // 128
cands = [i for i, (_, n, v, g) in enumerate(L) if 'conv_tc' in n and '<64>' in n and g.startswith('(%d,' % tiles)]
start = cands[-13] if len(cands) >= 13 else cands[0]
planes = [64, 128, 256, 512]; blocks = [3, 4, 6, 3]; strides = [1, 2, 2, 1]
h = h0; cin = 64; seq = []
for l in range(4):
    for b in range(blocks[l]):
        p = planes[l]; st = strides[l] if b == 0 else 1
        seq.append(('L%d.%d c1 %d->%d' % (l + 1, b, cin, p), 2 * B * h * h * cin * p))
        ho = h // st
        seq.append(('L%d.%d c2 3x3 s%d %d' % (l + 1, b, st, p), 2 * B * ho * ho * p * p * 9))
        if b == 0:
            seq.append(('L%d.%d ds %d->%d s%d' % (l + 1, b, cin, 4 * p, st), 2 * B * ho * ho * cin * 4 * p))
        seq.append(('L%d.%d c3 %d->%d' % (l + 1, b, p, 4 * p), 2 * B * ho * ho * p * 4 * p))
        cin = 4 * p; h = ho
# the stem (conv_tc<64> mode 1) may precede; skip launches until the first one whose grid matches layer1 c1
tc = [(n, v, g) for (_, n, v, g) in L[start:start + 90] if 'conv_tc' in n][:len(seq)]
tf = tt = 0
for (name, fl), (n, v, g) in zip(seq, tc):
    tf += fl; tt += v
    print('%-26s %8.1f us %7.1f TFLOP/s  grid %s' % (name, v / 1e3, fl / v / 1e3, g))
print('sum %.2f ms  %.1f TFLOP/s' % (tt / 1e6, tf / tt / 1e3))
