"""Probe: does torch's symmetric memory (CUDA P2P + NVLS multicast) rendezvous work on this box?  Run under torchrun."""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem

rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(int(os.environ['LOCAL_RANK']))
dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ['LOCAL_RANK'])))
t = symm_mem.empty(1 << 20, dtype=torch.float32, device='cuda')
h = symm_mem.rendezvous(t, dist.group.WORLD)
t.fill_(rank + 1)
dist.barrier()
torch.cuda.synchronize()
peer = h.get_buffer((rank + 1) % world, (16,), torch.float32)
print(f'rank {rank}: backend {symm_mem.get_backend(torch.device("cuda"))} multicast={h.has_multicast_support} mc_ptr={h.multicast_ptr:#x} '
      f'buffers={[hex(p) for p in h.buffer_ptrs]} signal_pads={[hex(p) for p in h.signal_pad_ptrs]} pad_size={h.signal_pad_size} '
      f'peer[0]={peer[0].item()}', flush=True)
dist.barrier()
dist.destroy_process_group()
