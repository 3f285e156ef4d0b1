#!/usr/bin/env python
"""bench.py -- training samples/s of the SASRec hot path (BASELINE.json metric) on N B200s, next to the CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

One "step" = one training iteration over one synthetic batch: gather -> SASRec encoder -> sampled-softmax loss ->
backward -> row-sparse table update + dense encoder update (unirec/facility/trainer.py:340-349).
Prints ONE JSON line (rank 0).  `value`: inputs resident in HBM.  `e2e`: same step through the public Trainer API with
pinned-host inputs copied in and the loss read back every step.  `roofline`: the fused gather/score/loss kernel
(HBM-bound), timed with CUDA events inside the timed region.  `cpu_baseline`: the oracle port of the reference step on
the host cores (bounded sample).  See DESIGN.md section "Measurement".
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # the configuration the BASELINE.json metric / north-star target is quoted on (fits one GPU: 5.1 GB table)
    'sasrec_d128_seq50_items10M_k1024_b1024': dict(model='SASRec', n_items=10_000_000, n_users=1_000_000, embedding_size=128,
                                                   hidden_size=128, n_layers=2, n_heads=2, inner_size=512, max_seq_len=50,
                                                   loss_type='softmax', K=1024, B=1024),
    # BASELINE.json configs[1]
    'sasrec_d128_seq50_items1M_k256_b1024': dict(model='SASRec', n_items=1_000_000, n_users=1_000_000, embedding_size=128,
                                                 hidden_size=128, n_layers=2, n_heads=2, inner_size=512, max_seq_len=50,
                                                 loss_type='softmax', K=256, B=1024),
    # BASELINE.json configs[2]
    'gru_d256_seq100_items5M_bpr5_b2048': dict(model='GRU', n_items=5_000_000, n_users=1_000_000, embedding_size=256,
                                               hidden_size=256, max_seq_len=100, loss_type='bpr', K=5, B=2048),
    # tiny: CI / CPU-baseline sanity
    'sasrec_tiny': dict(model='SASRec', n_items=20_000, n_users=1000, embedding_size=64, hidden_size=64, n_layers=1, n_heads=2,
                        inner_size=128, max_seq_len=20, loss_type='softmax', K=32, B=128),
}
DEFAULT_WORKLOAD = 'sasrec_d128_seq50_items10M_k1024_b1024'
COMMON = dict(dataset='example', exp_name='bench', train_file_format='user-item', hidden_dropout_prob=0.0,
              attn_dropout_prob=0.0, dropout_prob=0.0, scheduler='none', optimizer='adam', learning_rate=1e-3,
              hidden_act='swish', layer_norm_eps=1e-10, use_position_emb=1)


def synthetic_batch(w, seed):
    """uniform ids in [1, V-1], left-padded item_seq with length ~ U{1..L}, positive in column 0."""
    import torch
    g = torch.Generator().manual_seed(seed)
    V, U, B, K, L = w['n_items'], w['n_users'], w['B'], w['K'], w['max_seq_len']
    lens = torch.randint(1, L + 1, (B,), generator=g)
    seq = torch.randint(1, V, (B, L), generator=g, dtype=torch.int64)
    seq = torch.where(torch.arange(L)[None, :] >= (L - lens)[:, None], seq, torch.zeros_like(seq)).to(torch.int32)
    label = torch.zeros(B, 1 + K, dtype=torch.int32)
    label[:, 0] = 1
    return dict(user_id=torch.randint(1, U, (B,), generator=g), item_id=torch.randint(1, V, (B, 1 + K), generator=g),
                label=label, item_seq=seq, item_seq_len=lens)


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p['hbm_gbs']), float(p.get('bf16_tflops_sustained', p.get('bf16_tflops', 1590.0))), 'measured'
    return 6650.0, 1590.0, 'fallback'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith('active')})
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(sm)}


def run_reference(args, w, name):
    """--impl reference: the reference step (oracle port, dense grads + dense Adam) on the host cores."""
    import torch
    from oracle import cpu_baseline
    cfg = dict(COMMON)
    cfg.update({k: v for k, v in w.items() if k not in ('K', 'B')})
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    batches = [synthetic_batch(w, 100 + i) for i in range(2)]
    if w['model'] == 'MF':
        batches = [{k: b[k] for k in ('user_id', 'item_id', 'label')} for b in batches]
    sps, ms, threads = cpu_baseline.time_steps(w['model'], cfg, batches, args.steps, args.warmup)
    line = {'metric': 'training samples/sec', 'value': sps, 'unit': 'samples/s', 'impl': 'reference', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': name, 'batch_per_step': w['B'], 'optimizer': 'dense Adam (reference semantics)'},
            'cpu_baseline': {'value': sps, 'unit': 'samples/s', 'cores': threads, 'kind': 'port',
                             'sample': '%d timed steps of the full workload (B=%d) after %d warm-up' % (args.steps, w['B'], args.warmup)},
            'e2e': {'value': sps, 'unit': 'samples/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=None)
    ap.add_argument('--warmup', type=int, default=None)
    ap.add_argument('--workload', default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--gemm-precision', default='tf32x3', choices=['fp32', 'tf32x3', 'tf32'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-graph', action='store_true', help='keep Trainer.train_step eager (ncu launch lists)')
    ap.add_argument('--shard-p2p', type=int, default=1, help='row-sharded tables: 1 = peer-memory reads over NVLink, 0 = NCCL exchange')
    ap.add_argument('--overlap', type=int, default=0, help='overlap_table_update mode (0 off, 1 early link, 2 + early update)')
    args = ap.parse_args()
    sys.argv = sys.argv[:1]
    w = dict(WORKLOADS[args.workload])
    if args.impl == 'reference':
        args.steps = args.steps if args.steps is not None else 2
        args.warmup = args.warmup if args.warmup is not None else 1
        return run_reference(args, w, args.workload)
    args.steps = args.steps if args.steps is not None else 50
    args.warmup = args.warmup if args.warmup is not None else 5

    import torch
    import torch.distributed as dist
    from unirec_b200 import ops
    from unirec_b200.facility.accelerator import Accelerator
    from unirec_b200.facility.trainer import Trainer
    from unirec_b200.utils import argument_parser, general

    acc = Accelerator()
    world, rank, dev = acc.num_processes, acc.process_index, acc.device
    if world != args.gpus:
        raise SystemExit('--gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run for N>1)' % (args.gpus, world))
    B, K, L, d = w['B'], w['K'], w['max_seq_len'], w['embedding_size']
    cfg_args = dict(COMMON)
    cfg_args.update({k: v for k, v in w.items() if k not in ('K', 'B')})
    cfg_args.update(batch_size=B, n_sample_neg_train=K, gemm_precision=args.gemm_precision, epochs=1,
                    output_path=os.path.join(ROOT, 'gpurun_out', 'bench_ckpt'), table_shard_world=world,
                    overlap_table_update=args.overlap, cuda_graph=0 if args.no_graph else 1, shard_p2p=args.shard_p2p)
    cfg = argument_parser.parse_arguments(cfg_args, argv=[])
    cfg['device'] = dev
    general.init_seed(2022)
    model = general.get_class_instance(cfg['model'], 'unirec_b200/model')(cfg)
    trainer = Trainer(cfg, model, acc)

    n_pool = 8                      # distinct batches, rotated: no step re-reads the previous step's rows from L2
    host = [synthetic_batch(w, 1000 * rank + i) for i in range(n_pool)]
    if w['model'] == 'MF':
        host = [{k: b[k] for k in ('user_id', 'item_id', 'label')} for b in host]
    pinned = [{k: v.pin_memory() for k, v in b.items()} for b in host]
    resident = [{k: v.to(dev) for k, v in b.items()} for b in host]
    h2d_bytes = sum(v.numel() * v.element_size() for v in host[0].values())

    def barrier():
        if acc.distributed:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm (`value`) ----
    for i in range(args.warmup):
        trainer.train_step(resident[i % n_pool])
    barrier()
    ops.LAUNCH_COUNT = 0
    # sharded tables (N>1): the owner-side partial-softmax kernel is the row-gather kernel (same rows per rank per step)
    ops.TIMED_OP, ops.TIMED_EVENTS = ('ur_score_loss_fwd_bwd_f32' if world == 1 else 'ur_score_partial_f32'), []
    clocks = ClockSampler(dev.index or 0)
    clocks.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = trainer.train_step(resident[(args.warmup + i) % n_pool])
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    clk = clocks.stop()
    launches = ops.LAUNCH_COUNT
    kernel_ms = [a.elapsed_time(b) for a, b in ops.TIMED_EVENTS]
    ops.TIMED_OP, ops.TIMED_EVENTS = None, []
    final_loss = float(loss)

    # ---- same device-resident steps through the captured CUDA graph (Trainer.train_step's steady state) ----
    ms_graph = float('nan')
    if world == 1 and not args.no_graph:          # the row-sharded step (N > 1) is not graph-captured yet
        for i in range(3):
            trainer.train_step(resident[i % n_pool])
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for i in range(args.steps):
            trainer.train_step(resident[(args.warmup + i) % n_pool])
        g1.record()
        barrier()
        ms_graph = g0.elapsed_time(g1)

    # ---- end-to-end arm: pinned host inputs in, loss out, every step, through the public Trainer API ----
    # Trainer.device_batches is the loader-side API of the trainer: it copies batch i+1 host->device on a copy stream while
    # step i computes; every batch is copied from pinned host memory inside the timed region, every loss is read back.
    def host_stream(n):
        for i in range(n):
            yield pinned[i % n_pool]

    for batch in trainer.device_batches(host_stream(3)):
        float(trainer.train_step(batch))
    barrier()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for batch in trainer.device_batches(host_stream(args.steps)):
        float(trainer.train_step(batch))                   # device->host read of the step's loss
    t1.record()
    barrier()
    ms_e2e = t0.elapsed_time(t1)

    # ---- per-kernel breakdown (separate pass, events around every C-ABI call; not part of the timed numbers) ----
    ops.PROFILE = {}
    for i in range(3):
        trainer.train_step(resident[i % n_pool])
    torch.cuda.synchronize()
    breakdown = {k: sum(a.elapsed_time(b) for a, b in v) / 3.0 for k, v in ops.PROFILE.items()}
    ops.PROFILE = None

    t = torch.tensor([ms_total, ms_e2e, ms_graph], dtype=torch.float64, device=dev)
    if acc.distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e, ms_graph = float(t[0]), float(t[1]), float(t[2])
    if rank != 0:
        return
    hbm_peak, tc_peak, peak_kind = measured_peaks()
    samples = B * world * args.steps
    value = samples / (ms_total / 1e3)
    algo_bytes = B * (1 + K) * d * 4                     # target/negative rows read once by the fused score kernel
    # dram__bytes_read.sum + dram__bytes_write.sum per launch of this kernel from the committed `ncu --set full` capture
    # (profiles/r01d_full_captures.csv: 542.5 MB + 10.7 MB); only known for the default workload on one GPU
    traffic = 553.2e6 if (args.workload == DEFAULT_WORKLOAD and world == 1) else None
    k_ms = statistics.mean(kernel_ms) if kernel_ms else float('nan')
    achieved = algo_bytes / (k_ms / 1e3) / 1e9
    line = {
        'metric': 'training samples/sec', 'value': value, 'unit': 'samples/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_total / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'tf32' if args.gemm_precision == 'tf32' else 'f32', 'data': 'synthetic',
        'config': {'workload': args.workload, 'batch_per_gpu': B, 'global_batch': B * world, 'seq_len': L, 'n_items': w['n_items'],
                   'n_neg': K, 'optimizer': 'adam: row-sparse (lazy) on tables, flat dense on encoder', 'dropout': 0.0,
                   'gemm_precision': args.gemm_precision,
                   'encoder': 'live positions only (pack_sequences=1) and last layer on B rows (trim_last_layer=1): dead rows of the '
                              'reference computation are not computed, outputs and gradients identical (tests)',
                   'l2': 'inputs larger than L2: %.1f GB table, random rows, %d rotating batches'
                   % (w['n_items'] * d * 4 / 1e9, n_pool), 'parallelism': 'dp%d' % world, 'final_loss': final_loss},
        'clocks': clk,
        'e2e': {'value': samples / (ms_e2e / 1e3), 'unit': 'samples/s', 'h2d_bytes_per_step': h2d_bytes, 'd2h_bytes_per_step': 4,
                'ms_per_step': ms_e2e / args.steps},
        'gpu_launches': launches,
        'value_cuda_graph': None if ms_graph != ms_graph else {'value': samples / (ms_graph / 1e3), 'unit': 'samples/s', 'ms_per_step': ms_graph / args.steps,
                             'note': 'device-resident inputs, whole step replayed as one CUDA graph (the timed `value` region runs '
                                     'eagerly so that the roofline kernel can be bracketed by CUDA events)'},
        'roofline': {'kernel': 'score_loss_kernel (fused gather+dot+softmax+grad)' if world == 1 else
                     'score_partial_kernel (owner-side gather+dot+partial softmax, row-sharded table)', 'bound': 'hbm', 'achieved': achieved,
                     'peak': hbm_peak, 'unit': 'GB/s', 'frac': achieved / hbm_peak, 'traffic': traffic, 'peak_kind': peak_kind,
                     'algorithmic_bytes_per_launch': algo_bytes, 'kernel_ms': k_ms},
        'kernel_ms_per_step': {k: round(v, 4) for k, v in sorted(breakdown.items(), key=lambda kv: -kv[1])},
    }
    if not args.no_cpu_baseline and world == 1:          # the CPU baseline is reported at N=1 only
        from oracle import cpu_baseline
        cfgo = dict(COMMON)
        cfgo.update({k: v for k, v in w.items() if k not in ('K', 'B')})
        try:
            sps, ms, threads = cpu_baseline.time_steps(w['model'], cfgo, host[:2], steps=2, warmup=1)
            line['cpu_baseline'] = {'value': sps, 'unit': 'samples/s', 'cores': threads, 'kind': 'port', 'ms_per_step': ms,
                                    'sample': '2 timed steps of the full workload (B=%d, V=%d) after 1 warm-up, dense Adam' % (B, w['n_items'])}
        except (MemoryError, RuntimeError) as e:      # host RAM too small for dense grads + Adam state
            line['cpu_baseline'] = {'value': None, 'unit': 'samples/s', 'cores': os.cpu_count(), 'kind': 'port',
                                    'sample': 'failed: %s' % str(e)[:80]}
    print(json.dumps(line))


if __name__ == '__main__':
    main()
