import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import bench_data
from goofer_b200 import host
from oracle import dsp, resampler

i = 0
src, cli = bench_data.note_cli(i, "c3", 8)
f = bench_data.make_source(src)
env = dsp.decode_knots({"knot_vals_log": f["knot_vals_log"], "hz_knots": f["hz_knots"], "n_fft": 1024, "sr": 44100, "n_bins": 513})
feat = resampler.Features(env=env, mask=f["mask"], formants=f["formants"], sr=f["sr"], ylen=f["ylen"])
sf = host.SourceFeatures.from_knot_pack(f, f["mask"], f["formants"], f["sr"], f["ylen"])


def err(flags, pitch="A2"):
    c = list(cli); c[2] = flags; c[0] = pitch
    spec = resampler.NoteSpec.from_cli(*c)
    taps = {}
    ref = resampler.resample(feat, spec, lambda n, T: resampler.noise_for_note(spec, n, T, 20000 + 16 * i, 777 + i), taps=taps)
    b = host.Batch(); b.add_source(sf); b.add_note(host.NoteArgs.from_cli(0, c))
    ab = b.assemble(host.SeededNoise(20000 + 16 * i, 777 + i), taps=True)
    outs, tp = ab.render_host()
    e = np.abs(outs[0] - ref)
    blk = [float(e[k * 4410:(k + 1) * 4410].max()) for k in range(10)]
    return f"{e.max():.2e} @ {int(e.argmax())}  per-100ms: " + " ".join(f"{x:.0e}" for x in blk)


for fl in ("es20sg35", "es20sg100", "es60sg35", "es-40sg35", "sg35br-15", "sg35g-67", "sg35", "es20", "es20sg35P0", "es20sg35V50"):
    print(f"{fl:14s}", err(fl))
print("C4 es20sg35  ", err("es20sg35", "C4"))
