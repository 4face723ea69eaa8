import time, numpy as np, ctypes as C, sys
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from tbraymarcherplugin_b200 import _capi, synth, FMT_G8
from tbraymarcherplugin_b200.raymarch_utils import *
lib=_capi.load()
import os
flags=int(os.environ.get('TBRM_EXP','0'),0)
for n,view,steps in [(256,(512,512),256.),(512,(1920,1080),512.)]:
    res=URaymarchUtils.InitializeRaymarchResources((n,n,n),FMT_G8,bLightVolume32Bit=True)
    import torch
    d=torch.empty(n*n*n,dtype=torch.uint8,device='cuda')
    _capi.check(lib.tbrm_synth_volume_u8(0,1,(C.c_int32*3)(n,n,n),synth.PERLIN_SEED,C.c_void_p(d.data_ptr()),1))
    URaymarchUtils.SetDataVolumeDevice(res,d.data_ptr())
    URaymarchUtils.ColorCurveToTexture(res,synth.soft_ct_curve())
    URaymarchUtils.SetWindowingParameters(res,FWindowingParameters(0.45,0.5,True,False))
    w=synth.identity_world()
    URaymarchUtils.SetOptions(res,debug_flags=(flags,))
    for sync in (False, True):
      for it in range(3):
        ms=C.c_float()
        lib.tbrm_timer_begin(res.handle)
        URaymarchUtils.ClearResourceLightVolumes(res,0.0)
        for l in synth.LIGHTS[:2]: URaymarchUtils.AddDirLightToSingleVolume(res,l,True,w,bGPUSync=sync)
        lib.tbrm_timer_end(res.handle,C.byref(ms))
        print(n,'sweep reset 2 lights gpu_sync',sync,'ms',ms.value, flush=True)
    for li,l in enumerate(synth.LIGHTS):
        st=FSweepStats(); ms=C.c_float()
        lib.tbrm_timer_begin(res.handle)
        URaymarchUtils.AddDirLightToSingleVolume(res,l,True,w,bGPUSync=True,stats=st)
        lib.tbrm_timer_end(res.handle,C.byref(ms))
        print(n,'light',li,'faces',st.faces,'fused ms',ms.value, flush=True)
    cam=synth.benchmark_camera(*view)
    out=torch.empty(view[0]*view[1]*4,dtype=torch.float32,device='cuda')
    for it in range(3):
        ms=C.c_float()
        lib.tbrm_timer_begin(res.handle)
        _,st=URaymarchUtils.PerformWindowedLitRaymarch(res,cam,w,steps,device_out_ptr=out.data_ptr())
        lib.tbrm_timer_end(res.handle,C.byref(ms))
        print(n,'raymarch ms',ms.value,'steps',st,'Mray-steps/s',st/ms.value/1e3, flush=True)
    res.release()
