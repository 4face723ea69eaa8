import sys, ctypes as C
sys.path.insert(0,'.')
import torch
from tbraymarcherplugin_b200 import _capi, synth, FMT_G8
from tbraymarcherplugin_b200.raymarch_utils import *
lib=_capi.load()
n=int(sys.argv[1]) if len(sys.argv)>1 else 256
lights=[int(a) for a in sys.argv[2].split(',')] if len(sys.argv)>2 else [0,1]
res=URaymarchUtils.InitializeRaymarchResources((n,n,n),FMT_G8,bLightVolume32Bit=True)
d=torch.empty(n*n*n,dtype=torch.uint8,device='cuda')
_capi.check(lib.tbrm_synth_volume_u8(0,1,(C.c_int32*3)(n,n,n),synth.PERLIN_SEED,C.c_void_p(d.data_ptr()),1))
URaymarchUtils.SetDataVolumeDevice(res,d.data_ptr())
URaymarchUtils.ColorCurveToTexture(res,synth.soft_ct_curve())
URaymarchUtils.SetWindowingParameters(res,FWindowingParameters(0.45,0.5,True,False))
w=synth.identity_world()
for it in range(2):
    URaymarchUtils.ClearResourceLightVolumes(res,0.0)
    for l in lights: URaymarchUtils.AddDirLightToSingleVolume(res,synth.LIGHTS[l],True,w,bGPUSync=True)
    URaymarchUtils.FlushRenderingCommands(res)
print('done')
