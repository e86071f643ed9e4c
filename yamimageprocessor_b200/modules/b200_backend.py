"""Plugin package for the reference application: GPU-backed processing modules.

Add ``"yamimageprocessor_b200.modules"`` to ``AppConfiguration.plugin_packages``
(``core/app_core.py:55``); ``AppCore._discover_plugins`` (``:680-749``) imports this module and
calls ``register_module(app_core)``, which registers one ``ModuleBase`` subclass per operator.

* Identifiers and parameter names are the reference's own (``Grayscale``, ``NoiseReduction``,
  ``IntensityNormalization``, ... ; ``modules/preprocessing.py:46-134``) so ``pipeline_cache``
  signatures are unchanged; ops the reference lacks get new identifiers.
* ``pipeline_execution_metadata()`` says ``requires_gpu=True`` so ``PipelineManager.apply`` routes
  the step to a configured ``GpuExecutor`` (``processing/pipeline_manager.py:448-454``) ...
* ... and ``process()`` ITSELF runs on the GPU, because ``PipelineCache.compute``
  (``processing/pipeline_cache.py:379,496``) and ``run_enabled_stages`` (``ui/unified.py:558``)
  call ``step.apply`` directly and never see the executor.  There is no CPU path in either.
"""
from __future__ import annotations

from typing import Any, Dict, Mapping

import numpy as np

from ..host.params import MODULE_PARAMS, ParamSpec
from ..host.plugin import binding

ModuleBase, ModuleMetadata, ModuleStage, StepExecutionMetadata, _PipelineStep = binding()

_EXECUTOR = None


def _executor():
    global _EXECUTOR
    if _EXECUTOR is None:
        from ..host.executor import B200Executor

        _EXECUTOR = B200Executor()
    return _EXECUTOR


class _GpuModule(ModuleBase):
    """Common behaviour: parameter table, GPU execution hints, GPU ``process``."""

    IDENTIFIER = ""
    TITLE = ""
    STAGE = ModuleStage.PREPROCESSING
    DESCRIPTION = ""
    MENU = ("Pre-Processing",)

    def _build_metadata(self):
        return ModuleMetadata(
            identifier=self.IDENTIFIER,
            title=self.TITLE,
            stage=self.STAGE,
            description=self.DESCRIPTION,
            menu_path=self.MENU,
        )

    def parameter_metadata(self) -> Mapping[str, ParamSpec]:
        return MODULE_PARAMS.get(self.IDENTIFIER, {})

    def _load_parameter_metadata(self):  # the reference's ModuleBase calls this name
        return self.parameter_metadata()

    def pipeline_execution_metadata(self):
        return StepExecutionMetadata(supports_inplace=False, requires_gpu=True)

    def process(self, image: np.ndarray, **kwargs: Any) -> np.ndarray:
        params = self.sanitize_parameters(kwargs)
        ex = _executor()
        be = ex.backend
        out = ex.run_on_device(self.IDENTIFIER, be.to_device(np.asarray(image)), params)
        return be.to_host(out)


def _module(identifier: str, title: str, stage, description: str, menu=("Pre-Processing",)):
    return type(
        identifier.replace(" ", "") + "Module",
        (_GpuModule,),
        {"IDENTIFIER": identifier, "TITLE": title, "STAGE": stage, "DESCRIPTION": description, "MENU": menu,
         "__doc__": description, "__module__": __name__},
    )


GrayscaleModule = _module("Grayscale", "Toggle Greyscale", ModuleStage.PREPROCESSING,
                          "BGR -> gray (cv2.cvtColor BGR2GRAY, modules/preprocessing.py:52-55).")
BrightnessContrastModule = _module("BrightnessContrast", "Brightness / Contrast", ModuleStage.PREPROCESSING,
                                   "cv2.convertScaleAbs (modules/preprocessing.py:72-78).")
GammaCorrectionModule = _module("Gamma", "Gamma Correction", ModuleStage.PREPROCESSING,
                                "256-entry LUT (modules/preprocessing.py:95-102).")
IntensityNormalizationModule = _module("IntensityNormalization", "Intensity Normalization", ModuleStage.PREPROCESSING,
                                       "cv2.normalize NORM_MINMAX (modules/preprocessing.py:119-123).")
NoiseReductionModule = _module("NoiseReduction", "Noise Reduction", ModuleStage.PREPROCESSING,
                               "Gaussian / median denoising (modules/preprocessing.py:140-150).")
SharpenModule = _module("Sharpen", "Sharpen", ModuleStage.PREPROCESSING,
                        "Unsharp mask: addWeighted(img, 1+s, GaussianBlur(sigma 3), -s) (modules/preprocessing.py:167-171).")
SelectChannelModule = _module("SelectChannel", "Select Color Channel", ModuleStage.PREPROCESSING,
                              "Channel pick / two-channel mean (modules/preprocessing.py:188-209).")
ClaheModule = _module("CLAHE", "CLAHE", ModuleStage.PREPROCESSING,
                      "cv2.createCLAHE(clip_limit,(tile_grid_x,tile_grid_y)).apply (north_star op).")
BoxFilterModule = _module("BoxFilter", "Box Filter", ModuleStage.PREPROCESSING,
                          "cv2.blur(k,k), odd k (north_star op).")
HistogramEqualizationModule = _module("HistogramEqualization", "Histogram Equalization", ModuleStage.PREPROCESSING,
                                      "cv2.equalizeHist, uint8 (core/preprocessing.py:74-79).")
_SEG = ("Segmentation",)
GlobalThresholdModule = _module("Global", "Global Threshold", ModuleStage.SEGMENTATION,
                                "cv2.threshold BINARY (core/segmentation.py:140-143).", _SEG)
OtsuThresholdModule = _module("Otsu", "Otsu Threshold", ModuleStage.SEGMENTATION,
                              "cv2.threshold BINARY+OTSU (core/segmentation.py:145-148).", _SEG)
AdaptiveThresholdModule = _module("Adaptive", "Adaptive Threshold", ModuleStage.SEGMENTATION,
                                  "cv2.adaptiveThreshold GAUSSIAN_C (core/segmentation.py:91-94).", _SEG)
SobelModule = _module("Sobel", "Sobel", ModuleStage.SEGMENTATION,
                      "Sobel gradient magnitude (core/segmentation.py:150-155).", _SEG)
PrewittModule = _module("Prewitt", "Prewitt", ModuleStage.SEGMENTATION,
                        "Prewitt gradient magnitude (core/segmentation.py:157-164).", _SEG)
LaplacianModule = _module("Laplacian", "Laplacian", ModuleStage.SEGMENTATION,
                          "Absolute Laplacian (core/segmentation.py:166-169).", _SEG)
BorderRemovalModule = _module("Border Removal", "Border Removal", ModuleStage.SEGMENTATION,
                              "Clear a frame of border pixels (core/segmentation.py:316-325).", _SEG)
OpeningModule = _module("Opening", "Opening", ModuleStage.SEGMENTATION,
                        "cv2.morphologyEx OPEN (core/segmentation.py:264-275).", _SEG)
ClosingModule = _module("Closing", "Closing", ModuleStage.SEGMENTATION,
                        "cv2.morphologyEx CLOSE (core/segmentation.py:277-288).", _SEG)
DilationModule = _module("Dilation", "Dilation", ModuleStage.SEGMENTATION,
                         "cv2.dilate (core/segmentation.py:290-301).", _SEG)
ErosionModule = _module("Erosion", "Erosion", ModuleStage.SEGMENTATION,
                        "cv2.erode (core/segmentation.py:303-314).", _SEG)
ConnectedComponentsModule = _module("ConnectedComponents", "Connected Components", ModuleStage.SEGMENTATION,
                                    "8-connected labels in raster-first order (core/segmentation.py:108, core/extraction.py:60).", _SEG)
RegionLabelsModule = _module("RegionLabels", "Region Labels", ModuleStage.ANALYSIS,
                             "Otsu -> 8-connected labels (core/extraction.py:58-60,72-73); table: region_properties_data().", ("Extraction",))

class MosaicModule(_GpuModule):
    """Whole preprocess + segment pipeline over a LAZY tiled handle (BASELINE config 4).

    ``supports_tiled_input() -> True`` makes ``PipelineManager._apply_tiled`` hand the
    ``TiledPipelineImage`` itself to this step (``processing/pipeline_manager.py:412-416``,
    ``PipelineStep.apply`` ``:98-99``) instead of cutting halo-less tiles: rows are
    streamed from the handle (memmap slices) through a pinned ring into HBM and processed as row
    strips with over-fetched halos -- Gaussian -> CLAHE -> (Otsu) -> adaptive threshold -> open ->
    close -> connected components -- giving exactly the labels of the dense chain.  Under
    ``torch.distributed`` every process runs its own strips (LUT all-gather, histogram all-reduce,
    cross-strip label merge) and returns the label rows it owns.

    ``requires_gpu`` stays False here on purpose: ``_run_step`` densifies tiled input before calling a
    ``GpuExecutor`` (``:449-454``); this step's ``process`` is itself the GPU call."""

    IDENTIFIER = "Mosaic"
    TITLE = "Mosaic Preprocess + Segment"
    STAGE = ModuleStage.SEGMENTATION
    DESCRIPTION = "Row-strip sharded Gaussian -> CLAHE -> adaptive threshold -> open/close -> labels over a tiled handle."
    MENU = ("Segmentation",)

    def supports_tiled_input(self) -> bool:
        return True

    def pipeline_execution_metadata(self):
        return StepExecutionMetadata(supports_inplace=False, requires_gpu=False)

    def process(self, image, **kwargs: Any) -> np.ndarray:
        from ..host import ingest, mosaic

        p = self.sanitize_parameters(kwargs)
        mp = mosaic.MosaicParams(gauss_ksize=int(p["gauss_ksize"]), clip_limit=float(p["clip_limit"]),
                                 tile_grid=(int(p["tile_grid_x"]), int(p["tile_grid_y"])),
                                 block_size=int(p["block_size"]), C=float(p["C"]), morph_ksize=int(p["morph_ksize"]))
        be = _executor().backend
        if isinstance(image, np.ndarray) and image.ndim != 2:
            raise ValueError("Mosaic expects a single-channel (H, W) image or tiled handle")
        import os
        import time

        trace = os.environ.get("YAM_E2E_TRACE")
        t0 = time.perf_counter()
        results = mosaic.run_source(be, image, mp, strips_per_process=int(p["strips"]))
        t1 = time.perf_counter()
        self.last_results = results              # Otsu threshold / mask / CLAHE rows stay on the device for callers
        # Labels come back like every other step result (Backend.to_host): a fresh array backed by the
        # page-locked host pool, so the device -> host copy is one DMA per strip at PCIe speed with no
        # staging pass and no first-touch page faults on 16 GiB of new pageable memory.
        import torch

        rows = sum(int(r.labels.shape[0]) for r in results)
        out = be.pinned_empty((rows, int(results[0].labels.shape[1])), np.int32)
        y = 0
        for r in results:
            n = int(r.labels.shape[0])
            torch.from_numpy(out[y:y + n]).copy_(r.labels, non_blocking=True)
            y += n
        t2 = time.perf_counter()
        torch.cuda.current_stream(be.device).synchronize()
        if trace:
            import sys

            sys.stderr.write(f"[e2e] upload+compute {1e3 * (t1 - t0):.1f} ms, pinned alloc + enqueue {1e3 * (t2 - t1):.1f} ms, "
                             f"download wait {1e3 * (time.perf_counter() - t2):.1f} ms\n")
        return out


MODULE_CLASSES = (
    GrayscaleModule,
    BrightnessContrastModule,
    GammaCorrectionModule,
    IntensityNormalizationModule,
    NoiseReductionModule,
    SharpenModule,
    SelectChannelModule,
    ClaheModule,
    BoxFilterModule,
    HistogramEqualizationModule,
    GlobalThresholdModule,
    OtsuThresholdModule,
    AdaptiveThresholdModule,
    SobelModule,
    PrewittModule,
    LaplacianModule,
    BorderRemovalModule,
    OpeningModule,
    ClosingModule,
    DilationModule,
    ErosionModule,
    ConnectedComponentsModule,
    RegionLabelsModule,
    MosaicModule,
)


def register_module(app_core) -> None:
    """Plugin entry point (same contract as modules/preprocessing.py:270-274)."""
    for cls in MODULE_CLASSES:
        app_core.register_module(cls)


def region_properties_data(image: np.ndarray) -> Dict[str, np.ndarray]:
    """GPU counterpart of core/extraction.py:70-87: every column of the reference's table -- region_index,
    area, perimeter, centroid (row, col), eccentricity, solidity, extent, orientation -- plus bbox
    (half-open), area_convex and mean_intensity (extensions).  All sums, border-class counts and hull
    pixel counts are exact integers from the device; the float64 columns are formed from them on the host."""
    from ..host.steps import region_table

    ex = _executor()
    be = ex.backend
    t = be.to_device(np.asarray(image))
    gray = be.bgr2gray(t)
    labels = ex.run_on_device("RegionLabels", gray, {})
    return region_table(be, labels, gray if gray.dtype in _intensity_dtypes() else None)


def watershed_markers_data(image: np.ndarray, kernel_size: int = 3, opening_iterations: int = 2, dilation_iterations: int = 3,
                           distance_threshold_factor: float = 0.7) -> Dict[str, np.ndarray]:
    """GPU counterpart of the marker construction inside Detector.watershed_segmentation
    (core/segmentation.py:99-110) for a uint8 gray or BGR image: thresh, opening, sure_bg, dist (float32
    chamfer distance), sure_fg and the int32 markers image (0 unknown, 1 background, k + 1 markers).
    The flooding itself (cv2.watershed, :111) and the red boundary overlay the step returns (:112-114)
    are a sequential priority-queue algorithm plus UI drawing and stay with the reference: hand it
    ``markers`` (``cv2.watershed(image, markers)``) to finish the step from here."""
    ex = _executor()
    be = ex.backend
    gray = be.bgr2gray(be.to_device(np.asarray(image)))
    out = be.watershed_markers(gray, kernel_size, opening_iterations, dilation_iterations, distance_threshold_factor)
    return {k: be.to_host(v) for k, v in out.items()}


def hu_moments_data(image: np.ndarray) -> Dict[str, float]:
    """GPU counterpart of core/extraction.py:100-105: Otsu -> cv2.moments(mask) -> cv2.HuMoments,
    as ``{"hu_1": ..., "hu_7": ...}`` (the mask and its row power sums are formed on the device)."""
    from ..host import moments as M

    ex = _executor()
    be = ex.backend
    gray = be.bgr2gray(be.to_device(np.asarray(image)))
    mask = be.otsu_threshold(gray, 255)[1]
    rows = be.to_host(be.mask_row_moments(mask))
    hu = M.hu_moments(M.complete_moments(M.raw_moments_from_rows(rows)))
    return {f"hu_{i + 1}": float(v) for i, v in enumerate(hu)}


def histogram_data(image: np.ndarray) -> Dict[str, float]:
    """GPU counterpart of core/extraction.py:280-290 (uint8 gray levels): mean, variance, skewness
    and kurtosis of the 256-bin histogram, which is computed on the device."""
    from ..host import moments as M

    ex = _executor()
    be = ex.backend
    gray = be.bgr2gray(be.to_device(np.asarray(image)))
    if str(gray.dtype) != "torch.uint8":
        raise TypeError("histogram_data: the reference's 256-bin histogram is defined for uint8 images")
    return M.histogram_statistics(be.to_host(be.histogram(gray))[0])


def _intensity_dtypes():
    import torch

    return (torch.uint8, torch.uint16)


__all__ = [cls.__name__ for cls in MODULE_CLASSES] + ["MODULE_CLASSES", "register_module", "region_properties_data",
                                                       "hu_moments_data", "histogram_data", "watershed_markers_data"]
