import sys, os
from pathlib import Path; ROOT = Path(__file__).resolve().parents[1]; sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / 'tests'))
os.environ.setdefault("TBRM_TEST_SLAB_TIMEOUT_MS", "60000")
import numpy as np
if '--emu' in sys.argv:
    import emu_lib
    from tbraymarcherplugin_b200 import _capi
    _capi._lib = emu_lib.load()
import test_gpu_slab as T
from tbraymarcherplugin_b200 import synth
from tbraymarcherplugin_b200.raymarch_utils import URaymarchUtils
dims = tuple(int(a) for a in sys.argv[1].split(','))
nranks = int(sys.argv[2]); wname = sys.argv[3]
data = synth.perlin_ct_volume(dims)
world = T.WORLDS[wname]()
for li, light in enumerate(synth.LIGHTS):
    ref = T.unsharded.__wrapped__(data, [light], world) if hasattr(T.unsharded, '__wrapped__') else None
    res = T.make_res(data)
    URaymarchUtils.ClearResourceLightVolumes(res, 0.0)
    assert URaymarchUtils.AddDirLightToSingleVolume(res, light, True, world, bGPUSync=True)
    ref = URaymarchUtils.ReadLightVolume(res)
    ranks = T.virtual_ranks(data, nranks)
    for r, _, _ in ranks:
        URaymarchUtils.ClearResourceLightVolumes(r, 0.0)
    try:
        T.sharded_sweep(ranks, [light], world)
    except BaseException as e:
        print('light', li, 'FAILED', type(e).__name__, str(e)[:100]); continue
    got = T.merged(ranks)
    print('light', li, 'equal' if np.array_equal(got, ref) else f'DIFF {np.count_nonzero(got != ref)}', flush=True)
