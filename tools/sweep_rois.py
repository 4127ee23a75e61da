"""BASELINE config 5: RoIAlign / BA gather sweep over the number of RoIs (2 images, 256 channels,
7x7 bins, P2-P5), forward and backward, bf16 and fp32.  One JSON line per (size, kernel)."""
import json, os, subprocess, sys
here = os.path.dirname(os.path.abspath(__file__))
for rois in (128, 256, 512, 1024, 2048):
    out = subprocess.run([sys.executable, os.path.join(here, 'bench_kernels.py'), '--imgs', '2', '--rois', str(rois),
                          '--pos', str(rois // 4), '--iters', '10'], capture_output=True, text=True).stdout
    for ln in out.splitlines():
        if ln.startswith('{') and 'plan' not in ln and 'layout' not in ln:
            d = json.loads(ln)
            print(json.dumps(dict(rois_total=2 * rois, kernel=d['kernel'], dtype=d['dtype'].replace('torch.', ''), K=d.get('K'),
                                  ms=round(d['ms'], 4), alg_MB=round(d['alg_MB'], 1), GBs=round(d['GBs'], 1),
                                  frac=round(d['frac'], 3))), flush=True)
