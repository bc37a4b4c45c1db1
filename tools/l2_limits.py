import ctypes, torch, sys
sys.path.insert(0,'/root/repo')
from stac_speech_translation_b200 import _lib
a=ctypes.c_int64(); b=ctypes.c_int64()
torch.zeros(1,device='cuda')
_lib.lib().stac_l2_persist_limits(ctypes.byref(a), ctypes.byref(b)); print('max persisting L2', a.value/1e6, 'MB; max window', b.value/1e6, 'MB')
