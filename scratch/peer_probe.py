"""1-GPU probe: gg_peer_alloc + torch view over the raw allocation + barrier kernel at world=1 semantics."""
import ctypes, sys, torch
sys.path.insert(0, '/root/repo')
from graphgym_b200 import parallel
from graphgym_b200._lib import lib, check
L = lib()
ptr = ctypes.c_void_p(); h = ctypes.create_string_buffer(int(L.gg_peer_handle_bytes()))
check(L.gg_peer_alloc(1 << 20, ctypes.byref(ptr), h), 'alloc')
buf = parallel.PeerBuffer([ptr.value], 1 << 20, 0, torch.device('cuda', 0))
v = buf.view(256, 1024)
v.fill_(3.0)
torch.cuda.synchronize()
print('peer view ok', float(v.sum()), v.data_ptr() == ptr.value, v.device)
check(L.gg_peer_free(ptr), 'free')
