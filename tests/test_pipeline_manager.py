"""Driver semantics of the PipelineManager mirror, modelled on the reference's own tests
(tests/test_pipeline_manager.py:128-258, tests/test_processing_pipeline_manager_gpu.py:41-121,
tests/test_pipeline_streaming_large.py:111-133) and pinned by golden outputs of the reference."""
from __future__ import annotations

import logging
from pathlib import Path

import numpy as np
import pytest

from yamimageprocessor_b200.host.pipeline import PipelineManager, PipelineStep, StepExecutionMetadata
from yamimageprocessor_b200.host.plugin import ModuleBase, ModuleMetadata, ModuleStage
from yamimageprocessor_b200.host.tiles import TiledImageRecord, TiledPipelineImage, iter_tile_boxes

GOLD = Path(__file__).resolve().parent / "golden"


def add(image, *, value):
    return image + value


def mul(image, *, factor):
    return image * factor


def steps():
    return [PipelineStep("add", add, params={"value": 1.5}), PipelineStep("mul", mul, params={"factor": 2.0})]


def test_apply_matches_reference_golden():
    g = np.load(GOLD / "reference_outputs.npz")
    pm = PipelineManager(steps())
    assert np.array_equal(pm.apply(g["pm_in"]), g["pm_out"])
    assert np.array_equal(pm.apply(g["pm_stack_in"]), g["pm_stack_out"])  # slice-wise over planes


def test_apply_does_not_mutate_input_and_disabled_steps_are_skipped():
    img = np.ones((3, 3), np.float32)
    pm = PipelineManager(steps())
    out = pm.apply(img)
    assert np.array_equal(img, np.ones((3, 3), np.float32)) and np.allclose(out, 5.0)
    pm.set_step_enabled("mul", False)
    assert np.allclose(pm.apply(img), 2.5)
    assert pm.toggle_step("mul") is True


def test_undo_redo_and_history():
    pm = PipelineManager(steps())
    pm.push_state(image=np.zeros((2, 2)), cache_signature="a")
    pm.update_step_params("add", {"value": 4.0})
    assert pm.can_undo() and not pm.can_redo()
    state = pm.undo(current_cache_signature="b")
    assert state.cache_signature == "a" and pm.get_step("add").params["value"] == 1.5
    redo = pm.redo()
    assert redo.cache_signature == "b" and pm.get_step("add").params["value"] == 4.0
    assert pm.history_depth() == (1, 0)
    pm.reset()
    assert pm.history_depth() == (0, 0) and pm.get_step("add").params["value"] == 1.5


def test_ordering_and_editing_and_listeners():
    events = []
    pm = PipelineManager(steps())
    pm.add_change_listener(lambda e, m: events.append(e))
    pm.set_order(["mul"])
    assert [s.name for s in pm.steps] == ["mul", "add"]
    with pytest.raises(KeyError):
        pm.set_order(["nope"])
    pm.move_step(0, 1)
    pm.swap_steps(0, 1)
    pm.add_step(PipelineStep("third", add, params={"value": 0.0}), 1)
    assert pm.remove_step(1).name == "third"
    with pytest.raises(KeyError):
        pm.get_step("third")
    assert events == ["steps_reordered", "steps_reordered", "steps_swapped", "step_added", "step_removed"]
    clone = pm.clone()
    clone.update_step_params("add", {"value": 9.0}, replace=True)
    assert pm.get_step("add").params == {"value": 1.5}
    assert pm.to_dict()["steps"][0]["name"] == "mul"


def test_step_metadata_round_trip():
    step = PipelineStep("add", add, params={"value": 2.0},
                        execution=StepExecutionMetadata(supports_inplace=True, requires_gpu=True),
                        supports_tiled_input=True, stage=ModuleStage.SEGMENTATION)
    clone = step.clone()
    assert clone.execution == step.execution and clone.execution is not step.execution
    payload = step.to_dict()
    assert payload["execution"] == {"supports_inplace": True, "requires_gpu": True}
    assert payload["supports_tiled_input"] is True and payload["stage"] == "segmentation"
    back = PipelineStep.from_dict(payload, add)
    assert back.execution.requires_gpu and back.supports_tiled_input and back.stage is ModuleStage.SEGMENTATION


def test_inplace_and_none_results():
    def inplace(image, **_):
        image += 1
        return None

    img = np.zeros((2, 2), np.float32)
    step = PipelineStep("inc", inplace, execution=StepExecutionMetadata(supports_inplace=True))
    out = step.apply(img)
    assert out is img and np.all(img == 1)
    step2 = PipelineStep("copy", lambda image: image + 1, execution=StepExecutionMetadata(supports_inplace=True))
    out2 = step2.apply(img)
    assert out2 is img and np.all(img == 2)


class RecordingExecutor:
    def __init__(self):
        self.calls = []

    def execute(self, step, image):
        self.calls.append(step.name)
        return step.function(image, **step.params)


def test_gpu_executor_dispatch_and_cpu_warning(caplog):
    img = np.array([[0.0, 1.0], [2.0, 3.0]], np.float32)
    gpu_step = PipelineStep("gpu_add", add, params={"value": 1.0}, execution=StepExecutionMetadata(requires_gpu=True))
    ex = RecordingExecutor()
    pm = PipelineManager([gpu_step], gpu_executor=ex)
    assert np.allclose(pm.apply(img), img + 1) and ex.calls == ["gpu_add"]
    assert pm.clone()._gpu_executor is ex
    stack = np.zeros((3, 2, 2), np.float32)  # GPU branch receives the whole stack (no slicing)
    seen = []

    class ShapeExecutor:
        def execute(self, step, image):
            seen.append(image.shape)
            return None  # None -> input is returned

    pm.set_gpu_executor(ShapeExecutor())
    assert pm.apply(stack).shape == (3, 2, 2) and seen == [(3, 2, 2)]
    pm.set_gpu_executor(None)
    with caplog.at_level(logging.WARNING):
        out = pm.apply(img)
    assert "requires GPU execution" in caplog.text and np.allclose(out, img + 1)


def test_consecutive_gpu_steps_are_chained_once():
    class ChainExecutor(RecordingExecutor):
        def __init__(self):
            super().__init__()
            self.chains = []

        def execute_chain(self, steps, image):
            self.chains.append([s.name for s in steps])
            for s in steps:
                image = s.function(image, **s.params)
            return image

    gpu = StepExecutionMetadata(requires_gpu=True)
    pm = PipelineManager([PipelineStep("a", add, params={"value": 1.0}, execution=gpu),
                          PipelineStep("b", mul, params={"factor": 3.0}, execution=gpu),
                          PipelineStep("cpu", add, params={"value": 1.0}),
                          PipelineStep("c", add, params={"value": 2.0}, execution=gpu)],
                         gpu_executor=ChainExecutor())
    out = pm.apply(np.zeros((2, 2), np.float32))
    assert np.allclose(out, 6.0)
    assert pm._gpu_executor.chains == [["a", "b"]] and pm._gpu_executor.calls == ["c"]


def test_module_declares_gpu_requirement():
    class GpuModule(ModuleBase):
        def _build_metadata(self):
            return ModuleMetadata("gpu-module", "GPU Module", ModuleStage.PREPROCESSING)

        def pipeline_execution_metadata(self):
            return StepExecutionMetadata(requires_gpu=True)

        def process(self, image, **kwargs):
            return image

    step = GpuModule().create_pipeline_step()
    assert step.execution.requires_gpu and step.name == "gpu-module" and step.enabled is False


# ---- tiles -------------------------------------------------------------------------------------
def test_tile_boxes_are_disjoint_row_major_and_clipped():
    boxes = list(iter_tile_boxes(10, 7, (4, 3)))
    assert boxes[0] == (0, 0, 4, 3) and boxes[2] == (8, 0, 10, 3) and boxes[-1] == (8, 6, 10, 7)
    cover = np.zeros((7, 10), int)
    for l, t, r, b in boxes:
        cover[t:b, l:r] += 1
    assert np.all(cover == 1)
    assert list(iter_tile_boxes(5, 5, None)) == [(0, 0, 5, 5)]
    with pytest.raises(ValueError):
        list(iter_tile_boxes(5, 5, (0, 2)))


def test_tiled_apply_streams_npy_memmap(tmp_path):
    data = np.arange(16, dtype=np.float32).reshape(4, 4)
    path = tmp_path / "img.npy"
    np.save(path, data)
    rec = TiledImageRecord.from_npy(path)
    img = TiledPipelineImage(rec, tile_size=(2, 2))
    pm = PipelineManager([PipelineStep("add", add, params={"value": 1.0}), PipelineStep("mul", mul, params={"factor": 2.0})])
    out = pm.apply(img)
    assert np.array_equal(out, (data + 1) * 2)  # reference known answer (tests/test_pipeline_manager.py:237-251)
    assert np.array_equal(rec.read_region((1, 1, 3, 4)), data[1:4, 1:3])
    with pytest.raises(ValueError):
        rec.read_region((0, 0, 9, 9))
    assert img.infer_shape() == (4, 4) and img.dtype == np.float32
    rec.close()


def test_tiled_request_order_and_handle_passthrough():
    order = []

    class Synthetic:
        shape = (6, 8)
        size = None
        dtype = np.dtype(np.float32)

        def iter_tiles(self, tile_size):
            for box in iter_tile_boxes(8, 6, tile_size):
                order.append(box)
                l, t, r, b = box
                yy, xx = np.mgrid[t:b, l:r]
                yield box, (yy * 8 + xx).astype(np.float32)

        def to_array(self):
            raise AssertionError("streaming path must not densify")

    img = TiledPipelineImage(Synthetic(), tile_size=(4, 4))
    pm = PipelineManager([PipelineStep("s", lambda a: (a + 4) * 0.5)])
    out = pm.apply(img)
    yy, xx = np.mgrid[0:6, 0:8]
    assert np.array_equal(out, ((yy * 8 + xx) + 4) * 0.5)
    assert order == [(0, 0, 4, 4), (4, 0, 8, 4), (0, 4, 4, 6), (4, 4, 8, 6)]
    got = []
    pm2 = PipelineManager([PipelineStep("lazy", lambda h: got.append(type(h).__name__) or np.zeros((1, 1)),
                                        supports_tiled_input=True)])
    pm2.apply(img)
    assert got == ["TiledPipelineImage"]
