import sys, ctypes as C
sys.path.insert(0,'.')
import torch
from tbraymarcherplugin_b200 import _capi, synth, FMT_G8
from tbraymarcherplugin_b200.raymarch_utils import *
lib=_capi.load()
n=512
res=URaymarchUtils.InitializeRaymarchResources((n,n,n),FMT_G8,bLightVolume32Bit=True)
d=torch.empty(n*n*n,dtype=torch.uint8,device='cuda')
_capi.check(lib.tbrm_synth_volume_u8(0,1,(C.c_int32*3)(n,n,n),synth.PERLIN_SEED,C.c_void_p(d.data_ptr()),1))
URaymarchUtils.SetDataVolumeDevice(res,d.data_ptr())
URaymarchUtils.ColorCurveToTexture(res,synth.soft_ct_curve())
URaymarchUtils.SetWindowingParameters(res,FWindowingParameters(0.45,0.5,True,False))
w=synth.identity_world()
for l in synth.LIGHTS[:2]: URaymarchUtils.AddDirLightToSingleVolume(res,l,True,w,bGPUSync=True)
out=torch.empty(1920*1080*4,dtype=torch.float32,device='cuda')
cam=synth.benchmark_camera(1920,1080)
for it in range(3):
    URaymarchUtils.PerformWindowedLitRaymarch(res,cam,w,512.0,device_out_ptr=out.data_ptr())
print('done')
