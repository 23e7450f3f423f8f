#!/usr/bin/env python
"""bench.py — encode->decode throughput of the colour-modem hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--frames F] [--impl ours|reference]

Workload (BASELINE.json configs[1], the configuration the metric is quoted on): standard 625-line PAL, PAL-D
decoder, 720x576 frames.  A *step* is one pass of the hot path over one batch of F synthetic frames per GPU:
    composite = encode(rgb)      k_qam_encode
    rgb'      = decode(composite)   k_qam_bs_row (2 field-top rows per frame) + k_qam_rows<PALD> (all the filtering, one row
                                    per CTA) + k_qam_combine (elementwise pairing of neighbouring rows)
`value` is whole-job frames/s with the batch resident in HBM; `e2e` is the same metric through the public host
API (ImageModem.modulate_batch / demodulate_batch -> cm_encode_frames_host / cm_decode_frames_host) with pinned
HOST buffers, copies inside the timed region: two host threads keep both PCIe directions busy (batch i is demodulated
while batch i+1 is modulated); `e2e.sequential` is the same without the overlap.  Frames are sharded over ranks in contiguous ranges (rank r owns
absolute frames [r*F, (r+1)*F)); the path needs no inter-GPU traffic, so scaling is "weak".

`--impl reference` times the CPU restatement of the reference (oracle/, float64 numpy/scipy — the reference is
pure Python and cannot travel to the GPU box) on all host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

W, H = 720, 576
METRIC = 'enc->dec frames/s at 576i (PalDModem PAL 720x576)'
BYTES_PER_FRAME = 3 * W * H + W * H + W * H + 3 * W * H          # 3,317,760 B
DECODE_BYTES_PER_FRAME = W * H + 3 * W * H                        # composite in + RGB out of the decode kernel


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=40)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--frames', type=int, default=256, help='frames per GPU per step')
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--cpu-frames', type=int, default=0, help='frames in the CPU sample (0 = auto)')
    ap.add_argument('--no-extras', action='store_true', help='skip the short runs of the other BASELINE configs')
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on all host cores
# ------------------------------------------------------------------------------------------------------------
_worker_modem = None


def _cpu_init():
    global _worker_modem
    os.environ['OMP_NUM_THREADS'] = '1'
    import oracle
    _worker_modem = oracle.build(oracle.ModemSpec('pal_d', 'PAL', W, H))


def _cpu_frame(frame):
    from oracle import frame as oframe
    from color_modem_b200.synth import synth_frames_u8
    rgb = synth_frames_u8(1, H, W, first_frame=frame, seed=0)[0]
    t0 = time.perf_counter()
    comp = oframe.encode_frame_u8(_worker_modem, frame, rgb)
    out = oframe.decode_frame_u8(_worker_modem, frame, comp)
    return time.perf_counter() - t0, int(out[0, 0, 0])


def cpu_pool():
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        pass
    ctx = mp.get_context('fork')
    pool = ctx.Pool(cores, initializer=_cpu_init)
    pool.map(_cpu_frame, range(cores))          # untimed: spin up every worker, import scipy, design filters
    return pool, cores


def cpu_sample(pool, cores, nframes, first=0):
    t0 = time.perf_counter()
    per = pool.map(_cpu_frame, range(first, first + nframes), chunksize=max(1, nframes // cores))
    wall = time.perf_counter() - t0
    return nframes / wall, wall, sum(p[0] for p in per)


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    pool, cores = cpu_pool()
    nframes = args.cpu_frames or 2 * cores            # a whole number of frames per worker
    for _ in range(args.warmup):
        cpu_sample(pool, cores, cores)
    t0 = time.perf_counter()
    done = 0
    for s in range(args.steps):
        cpu_sample(pool, cores, nframes, first=s * nframes)
        done += nframes
    wall = time.perf_counter() - t0
    pool.close()
    fps = done / wall
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': fps, 'unit': 'frames/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * wall / max(args.steps, 1),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': 'PalDModem PAL 720x576 encode->decode (BASELINE configs[1]); CPU oracle port '
                               '(float64 numpy/scipy restatement of the reference, vectorised over lines), '
                               '%d frames per step over %d processes' % (nframes, cores)},
        'cpu_baseline': {'value': fps, 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
                         'sample': '%d steps x %d frames of 720x576 PAL-D encode->decode' % (args.steps, nframes)},
        'e2e': {'value': fps, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------
class ClockSampler(object):
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
             'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.QUERY,
                                          '--format=csv,noheader,nounits', '-lms', '50'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def stop(self, t_begin, t_end):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for t, ln in self.lines:
            if t < t_begin or t > t_end + 0.2:
                continue
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': smax, 'reasons': sorted(reasons),
                'samples': len(sm)}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


def run_ours(args):
    import torch
    import torch.distributed as dist
    from color_modem_b200 import _native as N
    from color_modem_b200.line import LineConfig
    from color_modem_b200.color.pal import PalDModem
    from color_modem_b200.image import ImageModem
    from color_modem_b200.synth import synth_frames_u8

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    cpu_line = None
    if world == 1:
        # CPU baseline beside the GPU number (N=1 only): oracle port on all host cores, bounded sample.  Taken before
        # CUDA is initialised so the worker processes can be forked safely.
        pool, cores = cpu_pool()
        nfr = args.cpu_frames or 2 * cores
        cfps, wall, cpu_s = cpu_sample(pool, cores, nfr)
        pool.close()
        pool.join()
        cpu_line = {'value': cfps, 'unit': 'frames/s', 'cores': cores, 'kind': 'port',
                    'sample': '%d frames of 720x576 PAL-D encode->decode through the float64 oracle port, '
                              '%d processes, %.1f s wall, %.1f s CPU' % (nfr, cores, wall, cpu_s)}
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    F = args.frames
    first_frame = rank * F                      # contiguous frame range of this rank

    modem = PalDModem(LineConfig((W, H)))
    # synthetic batch: 16 distinct frames tiled to F (content does not affect the data-independent kernels)
    base = synth_frames_u8(min(F, 16), H, W, first_frame=first_frame, seed=0)
    reps = -(-F // base.shape[0])
    host_rgb = torch.from_numpy(base).repeat(reps, 1, 1, 1)[:F].contiguous().pin_memory()
    rgb = host_rgb.to(dev)
    comp = torch.empty((F, H, W), dtype=torch.uint8, device=dev)
    out = torch.empty((F, H, W, 3), dtype=torch.uint8, device=dev)

    def step():
        modem.encode_frames(rgb, first_frame=first_frame, out=comp)
        modem.decode_frames(comp, first_frame=first_frame, out=out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    modem.timing(True)
    launches0 = N.launch_count()
    barrier()
    t_begin = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    t_end = time.perf_counter()
    launches = N.launch_count() - launches0
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    k_ms = {name: modem.timing_read(kid) for name, kid in
            (('encode', N.K_ENCODE), ('bandsplit_top_rows', N.K_BANDSPLIT), ('pald_rows', N.K_PALD),
             ('combine', N.K_DECODE_OTHER))}
    modem.timing(False)

    # ---- end-to-end through the public host API, pinned host buffers, copies inside the timed region --------
    img = ImageModem(modem)
    host_comp = torch.empty((F, H, W), dtype=torch.uint8).pin_memory()
    host_out = torch.empty((F, H, W, 3), dtype=torch.uint8).pin_memory()
    np_rgb, np_comp, np_out = host_rgb.numpy(), host_comp.numpy(), host_out.numpy()

    def e2e_step():
        img.modulate_batch(np_rgb, first_frame, out=np_comp)
        img.demodulate_batch(np_comp, first_frame, out=np_out)

    def timed(fn, n):
        barrier()
        t0 = time.perf_counter()
        fn(n)
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(2):
        e2e_step()
    e2e_seq_s = timed(lambda n: [e2e_step() for _ in range(n)], e2e_steps)

    # pipelined: a second modem handle (own streams and staging buffers) demodulates batch i while batch i+1 is being
    # modulated, so the H2D-heavy encode and the D2H-heavy decode share the full-duplex link.  Every timed step still
    # moves one batch of RGB in, its composite out and in again, and the decoded RGB out.
    from concurrent.futures import ThreadPoolExecutor
    img2 = ImageModem(PalDModem(LineConfig((W, H))))
    host_comp2 = torch.empty((F, H, W), dtype=torch.uint8).pin_memory()
    np_comp2 = host_comp2.numpy()
    comps = [np_comp, np_comp2]
    pool2 = ThreadPoolExecutor(max_workers=2)

    img2._modem._handle()                          # create the second native handle on this rank's device
    img.modulate_batch(np_rgb, first_frame, out=comps[0])              # pipeline fill (untimed)

    def e2e_pipelined(n):
        for i in range(n):
            a = pool2.submit(img.modulate_batch, np_rgb, first_frame, comps[(i + 1) & 1])
            b = pool2.submit(img2.demodulate_batch, comps[i & 1], first_frame, np_out)
            a.result()
            b.result()

    e2e_pipelined(2)
    e2e_s = timed(e2e_pipelined, e2e_steps)
    pool2.shutdown()
    checksum = int(np_out[0, :4, :4].sum())     # device->host read of the step's result

    # ---- the other BASELINE configs, briefly (rank 0, device-resident, CUDA events): context for the headline --------
    others = []
    if rank == 0 and not args.no_extras:
        from color_modem_b200 import comb
        from color_modem_b200.color import ntsc, secam, niir
        lc480 = LineConfig((720, 480))
        extra = [('Simple3DCombModem(NtscCombModem) NTSC 720x480 (BASELINE configs[2], north_star 480i target)',
                  lambda: comb.Simple3DCombModem(ntsc.NtscCombModem(lc480)), 480),
                 ('NtscModem NTSC 720x480 (BASELINE configs[0])', lambda: ntsc.NtscModem(lc480), 480),
                 ('ColorAveragingModem(SecamModem) SECAM 720x576 (BASELINE configs[3])',
                  lambda: comb.ColorAveragingModem(secam.SecamModem(LineConfig((W, H)))), H),
                 ('HueCorrectingNiirModem 720x576 (BASELINE configs[3])',
                  lambda: niir.HueCorrectingNiirModem(LineConfig((W, H))), H)]
        for name, make, hh in extra:
            mm = make()
            fr = 128
            xr = torch.from_numpy(synth_frames_u8(8, hh, W, seed=2)).repeat(fr // 8, 1, 1, 1).contiguous().to(dev)
            xc = mm.encode_frames(xr)
            xo = mm.decode_frames(xc)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                mm.encode_frames(xr, out=xc)
                mm.decode_frames(xc, out=xo)
            b.record()
            torch.cuda.synchronize()
            fps_x = 5 * fr / (a.elapsed_time(b) * 1e-3)
            bytes_x = 3 * W * hh + 2 * mm.composite_width * hh + 3 * mm.output_width * hh
            others.append({'workload': name, 'frames_per_s': fps_x, 'frames_per_step': fr,
                           'hbm_roofline_frac': fps_x * bytes_x / 1e9 / measured_peak_gbs()[0]})
            del mm, xr, xc, xo

    if rank == 0:
        total_frames = F * world * args.steps
        fps = total_frames / (ms * 1e-3)
        peak, peak_src = measured_peak_gbs()
        pald_ms, pald_n = k_ms['pald_rows']
        comb_ms, comb_n = k_ms['combine']
        per_launch_ms = pald_ms / max(pald_n, 1)
        frames_per_launch = F * args.steps / max(pald_n, 1)
        achieved = (frames_per_launch * DECODE_BYTES_PER_FRAME) / (per_launch_ms * 1e-3) / 1e9 if per_launch_ms > 0 else 0.0
        pair_ms = (pald_ms + comb_ms) / max(pald_n, 1)
        traffic = None
        try:
            with open(os.path.join(ROOT, 'profiles', 'r1_traffic.json')) as tf:
                traffic = json.load(tf)
        except Exception:
            pass
        step_ms = ms / args.steps
        shares = {k: (v[0] / max(v[1], 1)) * (v[1] / args.steps) / step_ms for k, v in k_ms.items()}
        line = {
            'metric': METRIC, 'value': fps, 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': step_ms, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'PalDModem PAL 720x576 encode->decode (BASELINE configs[1])',
                       'frames_per_gpu_per_step': F, 'sharding': 'contiguous frame ranges, no inter-GPU traffic',
                       'l2': 'no flush: each step streams %.0f MB per GPU (> 126 MB L2)' % (F * BYTES_PER_FRAME / 1e6)},
            'clocks': clocks,
            'e2e': {'value': F * world * e2e_steps / e2e_s, 'unit': 'frames/s',
                    'h2d_bytes_per_step': F * (3 * W * H + W * H), 'd2h_bytes_per_step': F * (W * H + 3 * W * H),
                    'steps': e2e_steps,
                    'api': 'ImageModem.modulate_batch / demodulate_batch on pinned host buffers; two host threads: '
                           'batch i is demodulated while batch i+1 is modulated',
                    'sequential': F * world * e2e_steps / e2e_seq_s,
                    'result_checksum': checksum},
            'gpu_launches': launches,
            'other_workloads': others,
            'roofline': {'bound': 'hbm', 'kernel': 'k_qam_rows<float, PALD>', 'achieved': achieved, 'peak': peak,
                         'unit': 'GB/s', 'frac': achieved / peak,
                         'traffic': (traffic['k_qam_rows']['dram_bytes_per_launch'] * frames_per_launch
                                     / traffic['k_qam_rows']['frames_per_launch']) if traffic else None,
                         'traffic_source': traffic['source'] if traffic else None,
                         'peak_source': peak_src,
                         'algorithmic_bytes_per_launch': frames_per_launch * DECODE_BYTES_PER_FRAME,
                         'frames_per_launch': frames_per_launch,
                         'avg_launch_ms': per_launch_ms,
                         'decode_pair': {'kernels': 'k_qam_rows<PALD> + k_qam_combine<PALD>', 'avg_ms': pair_ms,
                                         'achieved': (frames_per_launch * DECODE_BYTES_PER_FRAME) / (pair_ms * 1e-3) / 1e9
                                         if pair_ms > 0 else 0.0},
                         'kernel_share_of_step': shares,
                         'whole_chain_frac': fps / world * BYTES_PER_FRAME / 1e9 / peak,
                         'note': 'the decode chain is bound by instruction issue and FP32 latency, not by HBM: k_qam_rows runs the FMA pipe at 51 % and DRAM at 9 % of peak (profiles/r1_v8_pald_summary.md, DESIGN.md §5)'},
        }
        if cpu_line is not None:
            line['cpu_baseline'] = cpu_line
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
