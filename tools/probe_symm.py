"""Probe: does torch symmetric memory (CUDA VMM peer mappings between the ranks of one node) work on this box?
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/probe_symm.py
"""
import os

import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm_mem


def main():
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    t = symm_mem.empty(1 << 20, dtype=torch.float32, device=dev)
    t.fill_(float(rank + 1))
    hdl = symm_mem.rendezvous(t, dist.group.WORLD)
    print(f'rank {rank}: buffer_ptrs {[hex(p) for p in hdl.buffer_ptrs]} signal_pad_size {hdl.signal_pad_size} '
          f'multicast {hdl.has_multicast_support} mc_ptr {hex(hdl.multicast_ptr) if hdl.multicast_ptr else None}', flush=True)
    dist.barrier()
    peer = hdl.get_buffer((rank + 1) % world, (16,), torch.float32)
    torch.cuda.synchronize()
    print(f'rank {rank}: reads peer value {float(peer[0])} (expect {((rank + 1) % world) + 1})', flush=True)
    dist.barrier()
    peer[1] = 100.0 + rank                      # P2P store into the peer's buffer
    torch.cuda.synchronize()
    dist.barrier()
    print(f'rank {rank}: own [1] after peer store = {float(t[1])} (expect {100 + (rank - 1) % world})', flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
