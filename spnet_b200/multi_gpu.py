"""Data parallelism for the SPNet training step — the B200 replacement of the reference's
spnet/multi_gpu.py (make_parallel: single-process TF towers, per-GPU tf.slice of the batch,
outputs concatenated on the CPU, gradients summed implicitly through shared variables; disabled
in the reference at train_spnet.py:55).

Here: one process per GPU (torchrun), every rank holds a full replica and takes rows
[r*B/n, (r+1)*B/n) of the global batch (the reference's get_slice: shape[0]//parts, remainder
dropped, spnet/multi_gpu.py:49-54), BatchNorm statistics stay per replica exactly as in the tower
scheme (:61-79 calls the shared-weight model once per slice), and the gradients are averaged with
one NCCL all-reduce per step over NVLink 5 / NVSwitch, issued as buckets in the order backward
completes them: the Dense-head bucket (73 % of the bytes, offset 0 of the flat gradient buffer) and
the exit + middle-flow bucket are reduced on a side stream while backward is still running.
"""
import torch


def get_available_gpus():
    """Names in the reference's format ('/gpu:0', ...), spnet/multi_gpu.py:26-32."""
    if not torch.cuda.is_available():
        return []
    return ["/device:GPU:%d" % i for i in range(torch.cuda.device_count())]


def world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def batch_slice(n_rows, rank, parts):
    """Row range of this replica: the reference's get_slice (size = shape[0] // parts)."""
    size = n_rows // parts
    return rank * size, (rank + 1) * size


class GradAllReduce:
    """engine.grad_hook: average the flat gradient buffer over ranks in three buckets, in the order
    backward completes them: the Dense head (offset 0, 73 % of the bytes, final after the head's
    backward), the tail of the buffer (exit + middle flow, final after part A of the backbone
    backward) — both reduced on a side stream while backward continues — and the small remainder
    (stem, block 1, entry flow, residual convolutions) once backward is done."""

    def __init__(self, engine, group=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.world = dist.get_world_size(group)
        off, n, _ = engine.offsets["FinalOutput/kernel"]
        assert off == 0
        n_head = (n + 7) // 8 * 8
        t0 = engine.offsets[engine.tail_param_key][0] if engine.tail_param_key else engine.grads.numel()
        self.buckets = {"head": engine.grads[:n_head], "tail": engine.grads[t0:]}
        self.rest = engine.grads[n_head:t0]
        self.on_cuda = torch.device(engine.device).type == "cuda"
        if self.on_cuda:
            self.side = torch.cuda.Stream(device=engine.device)
            self.ready = {k: torch.cuda.Event() for k in self.buckets}
            self.done = {k: torch.cuda.Event() for k in self.buckets}
        self.in_flight = set()

    def bucket_ready(self, engine, which):
        """Called by the engine right after the gradients of bucket `which` are complete."""
        if not self.on_cuda or self.buckets[which].numel() == 0:
            return
        self.ready[which].record()
        with torch.cuda.stream(self.side):
            self.side.wait_event(self.ready[which])
            self.dist.all_reduce(self.buckets[which], group=self.group)
            self.done[which].record(self.side)
        self.in_flight.add(which)

    def __call__(self, engine):
        for which, buf in self.buckets.items():
            if which in self.in_flight:
                torch.cuda.current_stream().wait_event(self.done[which])
            elif buf.numel():
                self.dist.all_reduce(buf, group=self.group)
        self.in_flight.clear()
        if self.rest.numel():
            self.dist.all_reduce(self.rest, group=self.group)
        engine.optimizer_step(grad_scale=1.0 / self.world)
        engine.skip_default_optimizer = True


def attach_data_parallel(engine, group=None):
    hook = GradAllReduce(engine, group)
    engine.grad_hook = hook
    return hook


def make_parallel(model):
    """Keras-surface entry point (spnet/multi_gpu.py:35-88): returns the model wrapped for data
    parallelism when the process group has more than one rank, the model itself otherwise."""
    rank, n = world()
    if n < 2:
        return model
    model.parallel = True
    return model


def get_serial_part(model, parallel=True):
    """spnet/multi_gpu.py:15-23 — the replica itself (there is no wrapper layer to peel off)."""
    return getattr(model, "serial_model", model)
