"""Frame-range sharding across the GPUs of one box.

The reference carries no state from one frame to the next (every line memory is reset when ``frame`` changes),
so a frame sequence splits into contiguous ranges with **zero exchanged bytes**: rank k of G processes absolute
frames [k*N/G, (k+1)*N/G) and passes its absolute ``first_frame`` to the kernels (carrier phase, V-switch /
line-alternation parity and the SECAM inversion sequence depend on the absolute index).  There is no one-frame
halo — north_star's "NVLink halo for the 3D comb" has nothing to carry because the reference's "3D" combs use three
lines of the same field.  The only optional collective is gathering results on one rank.
"""


def frame_range(total_frames, rank, world_size):
    """Contiguous range [begin, end) of absolute frame indices owned by ``rank``."""
    if not (0 <= rank < world_size):
        raise ValueError('rank out of range')
    begin = (rank * total_frames) // world_size
    end = ((rank + 1) * total_frames) // world_size
    return begin, end


def process_sharded(total_frames, work, rank=None, world_size=None, gather=False):
    """Run ``work(first_frame, n_frames) -> tensor [n_frames, ...]`` on this rank's frame range.

    With ``gather=True`` the per-rank results are gathered (in frame order) on rank 0 through
    ``torch.distributed`` (NCCL over NVLink on GPUs, gloo in the CPU tests); other ranks get None.
    """
    import torch
    import torch.distributed as dist
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    begin, end = frame_range(total_frames, rank, world_size)
    out = work(begin, end - begin)
    if not gather or world_size == 1:
        return out
    # ranges differ by at most one frame: pad to the longest, gather, trim
    longest = max(frame_range(total_frames, r, world_size)[1] - frame_range(total_frames, r, world_size)[0]
                  for r in range(world_size))
    padded = torch.zeros((longest,) + tuple(out.shape[1:]), dtype=out.dtype, device=out.device)
    padded[:out.shape[0]] = out
    parts = [torch.empty_like(padded) for _ in range(world_size)] if rank == 0 else None
    dist.gather(padded, parts, dst=0)
    if rank != 0:
        return None
    keep = []
    for r in range(world_size):
        b, e = frame_range(total_frames, r, world_size)
        keep.append(parts[r][:e - b])
    return torch.cat(keep, dim=0)


def bind_to_gpu_numa_node(device_index):
    """Restrict this process to the CPUs of the NUMA node its GPU hangs off (Linux sysfs; a no-op when that cannot be
    determined).  Call it before allocating pinned staging buffers: first touch then places them on that node, so the
    host<->device copies of the host entry points do not cross the inter-socket link.  Returns the node or None."""
    import os
    try:
        import torch
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id      # torch >= 2.x: 'domain:bus:device.function'
    except Exception:                                                       # noqa: BLE001
        bus = None
    try:
        if not isinstance(bus, str):
            import ctypes
            rt = ctypes.CDLL('libcudart.so')
            buf = ctypes.create_string_buffer(32)
            if rt.cudaDeviceGetPCIBusId(buf, 32, int(device_index)) != 0:
                return None
            bus = buf.value.decode()
        bus = bus.lower()
        if len(bus.split(':')[0]) > 4:
            bus = bus[-12:]                                                  # sysfs uses a 4-digit domain
        with open('/sys/bus/pci/devices/%s/numa_node' % bus) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open('/sys/devices/system/node/node%d/cpulist' % node) as f:
            cpus = set()
            for part in f.read().strip().split(','):
                lo, _, hi = part.partition('-')
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:                                                       # noqa: BLE001
        return None
