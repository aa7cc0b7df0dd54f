"""FragmentStreams — several fragments in flight on ONE GPU.

A single fragment cannot fill a B200: the coarse levels launch 40-tile grids on 148 SMs, and every level ends in a
handful of 4-byte size read-backs (the sparsity is data dependent) during which the device idles.  Fragments of
different scenes are independent (the only shared state is the per-scene GRU volume,
models/gru_fusion.py:31-38), so the data-parallel unit is "one scene stream per CUDA stream": S replicas of NeuConNet
share one set of parameters but own their GRUFusion state, kernel-map caches and captured CUDA graphs; each replica
is driven by its own host thread on its own stream, so one stream's read-back bubble is filled by the others'
kernels.  The C ABI takes the stream explicitly and every ctypes / torch call releases the GIL, so plain threads work.

A scene's fragments must stay on one replica and in order (the GRU recurrence): route by `slot`.
"""
import queue
import sys
import threading

import torch

__all__ = ["replicate", "FragmentStreams"]


def replicate(net):
    """A second NeuConNet over the SAME parameter / buffer storage as `net` (no weight copy) with private per-scene
    state and caches."""
    from .neucon_network import NeuConNet
    rep = NeuConNet(net.cfg)
    # share the Parameter / buffer OBJECTS, not just their storage: every prepared-weight cache is keyed on
    # (data_ptr, _version), and a private Parameter would keep its own version counter -- a later load_state_dict or
    # optimizer step on `net` would then leave the replicas computing with stale weight slabs.
    src_mods = dict(net.named_modules())
    for mname, mod in rep.named_modules():
        src = src_mods[mname]
        for pname in list(mod._parameters):
            mod._parameters[pname] = src._parameters[pname]
        for bname in list(mod._buffers):
            mod._buffers[bname] = src._buffers[bname]
    rep.train(net.training)
    rep.with_panoptic_features = net.with_panoptic_features
    rep.with_panoptic = net.with_panoptic
    return rep


def _flatten(x):
    if isinstance(x, dict):
        for v in x.values():
            yield from _flatten(v)
    elif isinstance(x, (list, tuple)):
        for v in x:
            yield from _flatten(v)
    else:
        yield x


class _Future:
    def __init__(self):
        self._ev, self._val, self._exc = threading.Event(), None, None

    def set(self, val=None, exc=None):
        self._val, self._exc = val, exc
        self._ev.set()

    def result(self):
        self._ev.wait()
        if self._exc is not None:
            raise self._exc
        return self._val


SYNC_MODES = {"auto": 0, "spin": 1, "yield": 2, "blocking": 4}


def set_sync_mode(mode, device=None):
    """How host threads wait for the GPU in stream drains on `device` (ep_set_sync_mode): 'auto' (driver default: spin),
    'spin', 'yield' or 'blocking'.  Use 'yield' when streams x processes exceed the cores of the box (measured best at
    8 ranks x 8 streams on 32 CPUs; 'blocking' frees the cores completely but adds ~0.2 ms per drain)."""
    from . import _lib
    with torch.cuda.device(device if device is not None else torch.cuda.current_device()):
        _lib.check(_lib.lib().ep_set_sync_mode(SYNC_MODES[mode]), "ep_set_sync_mode")


class FragmentStreams:
    def __init__(self, net, n_streams, device=None, sync_mode=None):
        assert n_streams >= 1
        self.device = device if device is not None else next(net.parameters()).device
        if sync_mode is not None:
            set_sync_mode(sync_mode, self.device)
        self.nets = [net] + [replicate(net) for _ in range(n_streams - 1)]
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(n_streams)]
        self._queues = [queue.SimpleQueue() for _ in range(n_streams)]
        self._threads = []
        if sys.getswitchinterval() > 5e-4:
            sys.setswitchinterval(5e-4)   # host threads hand the GIL over at every launch; keep the fallback timer short
        for i in range(n_streams):
            t = threading.Thread(target=self._worker, args=(i,), daemon=True, name=f"eprecon-stream-{i}")
            t.start()
            self._threads.append(t)

    def __len__(self):
        return len(self.nets)

    def _worker(self, i):
        torch.cuda.set_device(self.device)
        q = self._queues[i]
        while True:
            job = q.get()
            if job is None:
                return
            fn, fut = job
            try:
                with torch.cuda.stream(self.streams[i]), torch.no_grad():
                    fut.set(fn(self.nets[i], self.streams[i]))
            except BaseException as e:  # surfaced by Future.result() on the submitting thread
                fut.set(exc=e)

    def submit(self, slot, fn, inputs=None):
        """Run fn(net_replica, cuda_stream) on worker `slot` (its stream is current); returns a future.  The worker's stream
        first waits for everything the SUBMITTING thread's current stream has enqueued so far (the producers of the
        fragment's tensors, e.g. a 2-D backbone); `inputs` (optional iterable of tensors allocated on other streams) are
        recorded on the worker stream so the caching allocator cannot recycle them while it still reads them."""
        fut = _Future()
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        stream = self.streams[slot]
        tensors = [t for t in _flatten(inputs) if torch.is_tensor(t) and t.is_cuda] if inputs is not None else []

        def job(net, st):
            st.wait_event(ev)
            for t in tensors:
                t.record_stream(st)
            return fn(net, st)
        self._queues[slot].put((job, fut))
        return fut

    def warm(self, fn):
        """Run fn once per replica, ONE AFTER THE OTHER, then synchronise the device: CUDA-graph captures and the shared
        lazily-built constant tables (kernel offsets, UMMA weight slabs) must exist before replicas run concurrently."""
        for i in range(len(self.nets)):
            self.submit(i, fn).result()
            torch.cuda.synchronize(self.device)

    def forward_many(self, fragments):
        """fragments: list of (features, features_backbone2d_occ_pano, inputs, outputs); fragment j runs on replica
        j % S.  Returns [(outputs, loss_dict)] in order (each replica's stream is synchronised)."""
        s = len(self.nets)

        def job(frag):
            def run(net, stream):
                res = net(*frag)
                stream.synchronize()
                return res
            return run
        futs = [self.submit(j % s, job(f), inputs=f[:3]) for j, f in enumerate(fragments)]
        return [f.result() for f in futs]

    def close(self):
        for q in self._queues:
            q.put(None)
        for t in self._threads:
            t.join(timeout=5)
