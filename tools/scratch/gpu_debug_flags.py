"""GPU bring-up: error of one workload note against the oracle, with flag groups removed one at a time."""
import sys, os, re
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import bench_data
from goofer_b200 import host
from oracle import dsp, resampler

workload, i = sys.argv[1], int(sys.argv[2])
src, cli = bench_data.note_cli(i, workload, 8)
f = bench_data.make_source(src)
env = dsp.decode_knots({"knot_vals_log": f["knot_vals_log"], "hz_knots": f["hz_knots"], "n_fft": 1024, "sr": 44100, "n_bins": 513})
feat = resampler.Features(env=env, mask=f["mask"], formants=f["formants"], sr=f["sr"], ylen=f["ylen"])
sf = host.SourceFeatures.from_knot_pack(f, f["mask"], f["formants"], f["sr"], f["ylen"])
flags = re.findall(r"[A-Za-z]+[+-]?\d+", cli[2])
print(cli)


def err(fl):
    c = list(cli); c[2] = "".join(fl)
    spec = resampler.NoteSpec.from_cli(*c)
    taps = {}
    ref = resampler.resample(feat, spec, lambda n, T: resampler.noise_for_note(spec, n, T, 20000 + 16 * i, 777 + i), taps=taps)
    b = host.Batch(); b.add_source(sf); b.add_note(host.NoteArgs.from_cli(0, c))
    ab = b.assemble(host.SeededNoise(20000 + 16 * i, 777 + i), taps=True)
    outs, tp = ab.render_host()
    e = np.abs(outs[0] - ref)
    return float(e.max()), int(e.argmax())


print("all flags", err(flags))
for k, fl in enumerate(flags):
    print("without", fl, err(flags[:k] + flags[k + 1:]))
for k, fl in enumerate(flags):
    print("only", fl, err([fl]))
