"""profiles/r02_rollout_error.md from the error-vs-step curves the long-rollout GPU tests write (gpurun_out/rollout_error_*.json)."""
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = ['# Round 2: error vs recurrent step, 32-step rollouts against the fp32 oracle (tests/test_gpu_model.py)', '',
       'B200, fp16 operands / fp32 accumulate, hidden state carried in fp16 between steps, B=1, synthetic count frames.',
       'Bars asserted at EVERY step: x_o max-abs <= 1e-2, |PSNR diff| <= 0.05 dB, x_o error <= 1 % of the learned residual,',
       'hidden-state error <= 1 % of the state\'s max-abs.  `resid` = max|x_o_ref - bilinear(f2)| (the learned part).', '']
for f in sorted(glob.glob(os.path.join(ROOT, 'gpurun_out', 'rollout_error_*.json'))):
    c = json.load(open(f))
    name = os.path.basename(f)[len('rollout_error_'):-5]
    out += ['## %s' % name, '', '| step | x_o max-abs | resid | x_o / resid | PSNR diff dB | hidden rel (max over states) |', '|---|---|---|---|---|---|']
    for r in c:
        out.append('| %d | %.2e | %.3f | %.5f | %.4f | %.5f |' % (r['step'], r['x_o_max_abs'], r['residual_max_abs'],
                                                              r['x_o_max_abs'] / r['residual_max_abs'], r['psnr_diff_db'], max(r['hidden_rel'])))
    first, last = c[:8], c[-8:]
    avg = lambda rows, k: sum(r[k] for r in rows) / len(rows)
    out += ['', 'max over %d steps: x_o %.2e, PSNR diff %.4f dB, hidden %.5f; mean x_o error steps 0-7 %.2e vs steps %d-%d %.2e (no drift).' % (
        len(c), max(r['x_o_max_abs'] for r in c), max(r['psnr_diff_db'] for r in c), max(max(r['hidden_rel']) for r in c),
        avg(first, 'x_o_max_abs'), len(c) - 8, len(c) - 1, avg(last, 'x_o_max_abs')), '']
open(os.path.join(ROOT, 'profiles', 'r02_rollout_error.md'), 'w').write('\n'.join(out))
print('\n'.join(out[-4:]))
