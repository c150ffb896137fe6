"""CPU: the scheduling logic of Network.forward_pipelined / _scene_loop (order of the per-scene host steps, slot reuse safety,
overlap depth) with the GPU steps replaced by recorders."""
import itertools

import pytest

from lib.nets.network import Network


class _Recorder:
    """Duck-typed `self` for Network.forward_pipelined: records (step, scene) and checks slot ownership."""

    def __init__(self):
        self.log = []
        self.slot_owner = {}   # slot index -> scene currently using it
        self.max_in_flight = 0
        self.in_flight = set()

    def _slot(self, i):
        return i

    def _stage_inputs(self, blobs, killing_inds, slot):
        assert slot not in self.slot_owner, f"slot {slot} restaged while scene {self.slot_owner.get(slot)} still owns it"
        self.slot_owner[slot] = blobs
        self.in_flight.add(blobs)
        self.max_in_flight = max(self.max_in_flight, len(self.in_flight))
        self.log.append(("stage", blobs))
        return dict(scene=blobs, slot=slot)

    def _run_static(self, h):
        self.log.append(("static", h["scene"]))
        return h

    def _launch_ragged(self, h):
        self.log.append(("ragged", h["scene"]))
        return h

    def _finalize(self, h):
        self.log.append(("final", h["scene"]))
        del self.slot_owner[h["slot"]]
        self.in_flight.discard(h["scene"])
        return {"scene": h["scene"]}


@pytest.mark.parametrize("n_scenes,n_static", list(itertools.product([0, 1, 2, 3, 5, 6, 7, 13, 40], [1, 2, 3, 4])))
def test_scene_loop_order_and_slot_safety(monkeypatch, n_scenes, n_static):
    monkeypatch.setenv("SIS3D_PIPE_STATIC", str(n_static))
    monkeypatch.delenv("SIS3D_PIPE_DEPTH", raising=False)
    rec = _Recorder()
    out = [(b, P["scene"]) for b, P in Network._scene_loop(rec, iter(range(n_scenes)))]
    assert out == [(i, i) for i in range(n_scenes)]          # every scene once, in order, with its own predictions
    assert not rec.slot_owner and not rec.in_flight          # everything drained
    assert rec.max_in_flight <= n_static + 3                 # staged + n_static replays + mask stage + read-back
    pos = {ev: k for k, ev in enumerate(rec.log)}
    for i in range(n_scenes):
        assert pos[("stage", i)] < pos[("static", i)] < pos[("ragged", i)] < pos[("final", i)]
        if i + 1 < n_scenes:  # inputs of the next scene are uploading before this scene's graph is launched
            assert pos[("stage", i + 1)] < pos[("static", i)]
        j = i + n_static - 1
        if j < n_scenes:      # n_static graph replays are queued before the host waits for the oldest one's detections
            assert pos[("static", j)] < pos[("ragged", i)]
