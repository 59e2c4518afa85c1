"""Result wrappers: ``Image`` and ``Video`` (reference ``image.py:12-277``).

Host-side only.  ``Image`` wraps the ndarray it is given without copying; frames
coming off the GPU are float64 ``(Nw, Nh)`` arrays like the reference's
``camera[:, :, 1]`` (``base.py:159``).
"""
import os.path
import pathlib

import numpy

__all__ = ["Image", "Video"]


class HostPlane(object):
    """A frame that came off the GPU as float32 (page-locked host memory) and stands for the float64
    ``(Nw, Nh)`` array scopyon hands out (``_epifm.py:1177``, ``base.py:159``).  ``(double)float`` is
    exact, so the float64 array is a pure function of the payload and is only materialised when
    somebody asks for it (``Image.as_array()``, ``numpy.asarray(plane)``): a movie that is scaled to
    8 bit, written to disk or merely counted never pays the 2 x 4 bytes per pixel of host memory
    traffic the widening costs.  ``widen(src, dst)`` is the engine's multi-threaded converter."""

    ndim = 2
    dtype = numpy.dtype(numpy.float64)

    def __init__(self, f32, widen=None, keep=None):
        assert f32.ndim == 2 and f32.dtype == numpy.float32
        self.f32 = f32
        self._widen = widen
        self._keep = keep          # owner of the payload's memory
        self._wide = None

    shape = property(lambda self: self.f32.shape)
    size = property(lambda self: self.f32.size)

    def widen(self):
        if self._wide is None:
            if self._widen is not None:
                self._wide = self._widen(self.f32)
            else:
                self._wide = self.f32.astype(numpy.float64)
        return self._wide

    def __array__(self, dtype=None, copy=None):
        wide = self.widen()
        return wide if dtype is None or numpy.dtype(dtype) == wide.dtype else wide.astype(dtype)


def _bytescale(data, cmin=None, cmax=None, low=None, high=None):
    """Linear map of [cmin, cmax] onto [low, high] as uint8 (``image.py:98-123``)."""
    cmin = data.min() if cmin is None else cmin
    cmax = data.max() if cmax is None else cmax
    low = 0 if low is None else low
    high = 255 if high is None else high
    span = cmax - cmin
    if span == 0.0:
        return numpy.ones(data.shape, dtype=numpy.uint8) * low
    out = (data - cmin) * (float(high - low) / span) + low
    return (out.clip(low, high) + 0.5).astype(numpy.uint8)


class Image(object):

    def __init__(self, data):
        assert data.ndim == 2 or (data.ndim == 3 and data.shape[2] == 3)
        self.__payload = data      # an ndarray, or a HostPlane until the float64 array is asked for

    @property
    def __data(self):
        return self.as_array()

    @staticmethod
    def load(file):
        assert isinstance(file, (str, pathlib.PurePath))
        filename = str(file)
        ext = os.path.splitext(filename)[1].lower()
        if ext == '.npy':
            return Image(numpy.load(filename))
        if ext == '.csv':
            return Image(numpy.loadtxt(filename))
        import PIL.Image
        return Image(numpy.asarray(PIL.Image.open(filename)))

    @staticmethod
    def RGB(red=None, green=None, blue=None):
        """Stack up to three same-shape channels into an ``(H, W, 3)`` image
        (``image.py:38-68``) -- the two-colour assembly of ``examples/twocolor.py``."""
        channels = (red, green, blue)
        given = [c for c in channels if c is not None]
        assert len(given) > 0
        shape, dtype = given[0].shape, given[0].dtype
        assert all(c.shape == shape and c.dtype == dtype for c in given)
        rgb = numpy.zeros((shape[0], shape[1], 3), dtype=dtype)
        for i, channel in enumerate(channels):
            if channel is None:
                continue
            assert isinstance(channel, (Image, numpy.ndarray))
            rgb[:, :, i] = channel.as_array() if isinstance(channel, Image) else channel
        return Image(rgb)

    def as_array(self, dtype=None):
        """The image as an ndarray (the reference's ``Image.as_array()``, ``image.py:70-71``).

        Frames formed on the GPU in float32 are widened to the float64 array of the reference on the
        first call (exact; the array is then kept).  ``as_array(numpy.float32)`` -- an extension --
        returns the float32 payload itself without widening when the frame still holds one."""
        data = self.__payload
        if isinstance(data, HostPlane):
            if dtype is not None and numpy.dtype(dtype) == numpy.float32:
                return data.f32
            data = self.__payload = data.widen()       # the float32 payload (pinned memory) is released
        if dtype is not None and numpy.dtype(dtype) != data.dtype:
            return data.astype(dtype)
        return data

    dtype = property(lambda self: self.__payload.dtype)
    ndim = property(lambda self: self.__payload.ndim)
    size = property(lambda self: self.__payload.size)
    shape = property(lambda self: self.__payload.shape)

    def as_8bit(self, cmin=None, cmax=None, low=None, high=None):
        if self.dtype == numpy.uint8:
            return Image(self.__data.copy())
        if self.ndim == 2:
            return Image(_bytescale(self.__data, cmin, cmax, low, high))
        data = numpy.zeros(self.shape, dtype=numpy.uint8)
        for i in range(3):
            data[:, :, i] = _bytescale(self.__data[:, :, i], cmin, cmax, low, high)
        return Image(data)

    def save(self, filename, **kwargs):
        """``.npy`` / ``.csv`` store the raw array; anything else is written as an
        8-bit picture through pillow (optionally with ``shapes`` boxes drawn)."""
        assert isinstance(filename, (str, pathlib.PurePath))
        filename = str(filename)
        ext = os.path.splitext(filename)[1].lower()
        if ext == '.npy':
            assert len(kwargs) == 0
            numpy.save(filename, self.__data)
        elif ext == '.csv':
            assert len(kwargs) == 0
            numpy.savetxt(filename, self.__data)
        else:
            self.savefig(filename, self.as_8bit().as_array(), **kwargs)

    @staticmethod
    def savefig(filename, img, shapes=None):
        import PIL.Image
        from PIL import ImageDraw
        picture = PIL.Image.fromarray(img)
        if shapes is not None:
            picture = picture.convert('RGB')
            draw = ImageDraw.Draw(picture)
            for shape in shapes:
                row, column, half = shape['x'], shape['y'], shape['sigma']
                draw.rectangle([(column - half, row - half), (column + half, row + half)],
                               outline=shape['color'], width=1)
        picture.save(filename)

    def show(self, **kwargs):
        try:
            import matplotlib.pyplot as plt
        except ImportError as exc:  # display is outside the accelerated path
            raise ImportError("Image.show() needs matplotlib") from exc
        plt.imshow(self.as_8bit().as_array(), interpolation='none', cmap='gray')
        plt.show()


class Video:

    @staticmethod
    def save(filename, imgs, interval=100, dpi=None, cmin=None, cmax=None, low=None, high=None):
        """Write a movie.  ``.npy`` stores the 8-bit frame stack directly; ``.gif`` goes
        through pillow; other containers need matplotlib + ffmpeg like the reference
        (``image.py:240-277``)."""
        if cmin is None:
            cmin = min(img.as_array().min() for img in imgs)
        if cmax is None:
            cmax = max(img.as_array().max() for img in imgs)
        frames = [img.as_8bit(cmin=cmin, cmax=cmax, low=low, high=high).as_array() for img in imgs]
        ext = os.path.splitext(str(filename))[1].lower()
        if ext == '.npy':
            numpy.save(str(filename), numpy.stack(frames))
            return
        if ext == '.gif':
            import PIL.Image
            pics = [PIL.Image.fromarray(f) for f in frames]
            pics[0].save(str(filename), save_all=True, append_images=pics[1:], duration=interval, loop=0)
            return
        import matplotlib.pyplot as plt
        from matplotlib.animation import FuncAnimation
        plt.ioff()
        fig, ax = plt.subplots(1, figsize=(1, 1))
        fig.subplots_adjust(0, 0, 1, 1)
        ax.axis("off")
        artist = ax.imshow(frames[0], cmap='gray', vmin=0, vmax=255)

        def animate(k):
            artist.set_array(frames[k])
            return artist
        FuncAnimation(fig, animate, numpy.arange(len(frames)), interval=interval).save(
            str(filename), dpi=dpi or max(frames[0].shape))
        plt.clf()
        plt.ion()
