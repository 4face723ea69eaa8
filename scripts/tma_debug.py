import sys, numpy as np
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from tbraymarcherplugin_b200 import synth, FMT_G8
from tbraymarcherplugin_b200.raymarch_utils import *
import oracle
dims=tuple(int(a) for a in sys.argv[1:4]) if len(sys.argv)>3 else (32,32,32)
li=int(sys.argv[4]) if len(sys.argv)>4 else 3
data=synth.perlin_ct_volume(dims)
res=URaymarchUtils.InitializeRaymarchResources(dims,FMT_G8,bLightVolume32Bit=True)
URaymarchUtils.SetDataVolume(res,data); URaymarchUtils.ColorCurveToTexture(res,synth.soft_ct_curve())
w=FWindowingParameters(0.45,0.5,True,False); URaymarchUtils.SetWindowingParameters(res,w)
URaymarchUtils.SetOptions(res,sweep_impl=2)
ora=oracle.OracleVolume(data,oracle.prepare_tf(synth.soft_ct_curve()),w)
st=FSweepStats()
URaymarchUtils.AddDirLightToSingleVolume(res,synth.LIGHTS[li],True,synth.identity_world(),bGPUSync=True,stats=st)
print(st)
g=URaymarchUtils.ReadLightVolume(res)
ora.add_dir_light(synth.LIGHTS[li],True,synth.identity_world())
d=np.abs(g-ora.light); print('max diff',d.max(),'nonzero',np.count_nonzero(d),'of',d.size)
if d.max()>0:
    idx=np.argwhere(d>0); print(idx[:10], idx.min(0), idx.max(0))
