"""Batched drop-in for the reference's pose iterators
(python-examples/position-estimation-toy-experiment/compoundRayIterators.py:27-151).

Same classes, constructor arguments and yielded tuples as the reference's ``RandomCubeIterator`` /
``UniformCubeIterator``; the difference is inside: instead of one ``setCameraPosition`` +
``renderFrame`` + ``getFramePointer`` round trip per item, blocks of poses are rendered by
``crRenderPoseBatch`` (several frames per kernel launch, one device->host copy per block) and handed
out one by one.  Item k of the iterator is frame k of every RNG stream, so the images are byte-identical to
the reference loop's: a block that is abandoned part-way (``iter()`` called again, an epoch that ends inside a
block) rewinds the streams to the number of items actually handed out (``crSetFirstFrame``) before the next
block is rendered.  ``RandomCubeIterator`` draws positions from numpy's global RNG exactly as the reference does
(``np.random.random(3)`` per item, in order) but ``blockSize`` items AHEAD of the item it returns: code that
interleaves its own ``np.random`` calls with ``next()`` sees a different global sequence than with the reference
loop (use ``blockSize=1`` there).
"""
from __future__ import annotations

import math
import os

import numpy as np

import eye_renderer as er


class CompoundRayIterator:
    def __init__(self, eyeFilepath, debug=False, debugPano=True, transform=None, resultNormalisationData=None,
                 scenePath="sim-environment/env_2.gltf", cameraName="compound-cam", samples=1000, blockSize=64,
                 libPath=None, device=None):
        self.debug = debug
        self.eyeRenderer = er.load_library(libPath, device=device)
        self.eyeRenderer.setVerbosity(bool(debug))
        self.eyeRenderer.loadGlTFscene(os.fsencode(scenePath))
        self.eyeRenderer.gotoCameraByName(cameraName.encode())
        eyeConfig = er.readEyeFile(eyeFilepath)
        er.setOmmatidiaFromOmmatidiumList(self.eyeRenderer, eyeConfig)
        self.eyeRenderer.setCurrentEyeShaderName(b"single_dimension_fast")
        self.ommatidia = len(eyeConfig)
        er.setRenderSize(self.eyeRenderer, self.ommatidia, 1)
        self.eyeRenderer.setCurrentEyeSamplesPerOmmatidium(int(samples))
        pose = np.zeros(12, np.float32)
        self.eyeRenderer.crDebugCopyCameraPose(pose.ctypes.data)
        self._axes = pose[3:].copy()
        self.resultNormalisationData = resultNormalisationData
        self.tf = transform
        self.blockSize = int(blockSize)
        self._rows = None
        self._positions = None
        self._cursor = 0
        self._consumed = 0          # items handed out since the streams were initialised = the frame the next item must be

    def __iter__(self):
        return self

    def _drop_block(self):
        """Forget a block whose tail was rendered but never handed out: the streams go back to frame `_consumed`."""
        if self._rows is not None and self._cursor < len(self._rows):
            self.eyeRenderer.crSetFirstFrame(int(self._consumed))
        self._rows = None

    def _render_block(self, positions):
        poses = er.make_poses(positions, x=self._axes[0:3], y=self._axes[3:6], z=self._axes[6:9])
        rows, _ = er.renderPoseBatch(self.eyeRenderer, poses)
        self._rows = rows.reshape(len(positions), 1, self.ommatidia, 4)      # (H=1, W=N, 4) like getFramePointer()
        self._positions = positions
        self._cursor = 0

    def stop(self):
        self.eyeRenderer.stop()


class RandomCubeIterator(CompoundRayIterator):
    """Camera at a uniformly random position inside a cube of side `cubeSize` per item."""

    def __init__(self, eyeFilepath, debug=False, cubeSize=50, debugPano=True, transform=None, resultNormalisationData=None, **kw):
        super().__init__(eyeFilepath, debug, debugPano, transform=transform, resultNormalisationData=resultNormalisationData, **kw)
        self.cubeSize = cubeSize

    def __next__(self):
        import torch
        if self._rows is None or self._cursor >= len(self._rows):
            rel = (np.random.random((self.blockSize, 3)) * 2 - 1) * (self.cubeSize / 2)   # same draws as B x random(3)
            self._render_block(rel)
        i = self._cursor
        self._cursor += 1
        self._consumed += 1
        image = np.copy(self._rows[i][:, :, :3])
        return torch.from_numpy(image.astype(np.dtype("f"))), torch.from_numpy(self._positions[i].astype(np.dtype("f")))


class UniformCubeIterator(CompoundRayIterator):
    """Camera on a regular samplingSize^3 lattice inside the cube, x fastest."""

    def __init__(self, eyeFilepath, debug=False, cubeSize=50, samplingSize=100, debugPano=True, transform=None,
                 resultNormalisationData=None, **kw):
        super().__init__(eyeFilepath, debug, debugPano, transform=transform, resultNormalisationData=resultNormalisationData, **kw)
        self.cubeSize = cubeSize
        self.samplingSize = samplingSize

    def __iter__(self):
        self.sampleID = 0
        self.sampleGap = self.cubeSize / (self.samplingSize + 1)
        self.startPos = np.ones(3) * (-(self.samplingSize * self.sampleGap) / 2)
        self._drop_block()
        return self

    def _coord(self, sid):
        n = self.samplingSize
        z = math.floor(sid / (n ** 2))
        y = math.floor((sid - z * (n ** 2)) / n)
        x = sid - z * (n ** 2) - y * n
        return np.asarray([x, y, z], dtype=np.int32)

    def __next__(self):
        import torch
        total = self.samplingSize ** 3
        if self._rows is None or self._cursor >= len(self._rows):
            ids = [(self.sampleID + k) % total for k in range(self.blockSize)]
            self._coords = [self._coord(s) for s in ids]
            pos = np.stack([c * np.ones(3) * self.sampleGap + self.startPos for c in self._coords])
            self._render_block(pos)
        i = self._cursor
        self._cursor += 1
        self._consumed += 1
        self.sampleID = (self.sampleID + 1) % total
        coord = self._coords[i]
        samplingPos = self._positions[i]
        image = np.copy(self._rows[i][:, :, 0])
        imageOut = torch.from_numpy(image.astype(np.dtype("f")))
        vectorOut = torch.from_numpy(samplingPos.astype(np.dtype("f")))
        if self.tf is not None:
            imageOut = self.tf(image.astype(np.dtype("f")))
        if self.resultNormalisationData is not None:
            vectorOut = (vectorOut - self.resultNormalisationData["means"]) / self.resultNormalisationData["stds"]
        return imageOut, vectorOut, coord

    def getSamplingSize(self):
        return self.samplingSize

    def getTotalSamplePoints(self):
        return self.samplingSize ** 3
