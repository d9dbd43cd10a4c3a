"""Ball initializers -- mirror of lsml/initializer/provided/ball.py:6-115.

``initialize`` (host, numpy) keeps the reference semantics for drop-in use; ``device_ball``
gives the evolution engine the parameters to build ``u0`` directly in HBM with
lsml_ball_init (no host grid, no upload) -- same float64 expression, bit-identical.
"""
import numpy

from ..initializer_base import InitializerBase


class BallInitializer(InitializerBase):
    """ Initialize the zero level set to a ball of fixed radius """

    def __init__(self, radius=10, location=None):
        self.radius = radius
        self.location = location

    def device_ball(self, img_shape, dx, first_value=None):
        """(centre in physical units, radius) for the on-device initializer."""
        if self.location is not None and len(self.location) != len(img_shape):
            msg = '`location` is len {} but should be {}'
            raise ValueError(msg.format(len(self.location), len(img_shape)))
        location = (0.5 * numpy.array(img_shape) if self.location is None
                    else numpy.asarray(self.location, dtype=float))
        return location * dx, float(self.radius)

    def initialize(self, img, dx, seed):
        if self.location is not None and len(self.location) != img.ndim:
            msg = '`location` is len {} but should be {}'
            raise ValueError(msg.format(len(self.location), img.ndim))
        if self.location is None:
            location = 0.5 * numpy.array(img.shape)
        else:
            location = self.location
        slices = (slice(None),) + tuple(None for _ in range(img.ndim))
        indices = numpy.indices(img.shape, dtype=float)
        indices *= dx[slices]
        indices -= (location * dx)[slices]
        return (self.radius - numpy.sqrt((indices**2).sum(axis=0))) > 0


class RandomBallInitializer(InitializerBase):
    """ Ball with random centre and radius, seeded from the image (reference :34-115) """

    def __init__(self, randomize_center=True, random_state=None):
        if random_state is None:
            random_state = numpy.random.RandomState()
        self.random_state = random_state
        self.randomize_center = randomize_center

    def _get_seed_value_from_image(self, img):
        img_str = "{:.4f}".format(img.ravel()[0])
        _, decimal_str = img_str.split(".")
        return int(decimal_str)

    def _draw(self, img_shape, dx, first_value):
        """centre (physical units) and radius, consuming the RNG exactly like the reference."""
        img_str = "{:.4f}".format(first_value)
        seed_value = int(img_str.split(".")[1])
        state = self.random_state.get_state()
        self.random_state.seed(seed_value)
        min_dim = min(dx * img_shape)
        radius = self.random_state.uniform(low=0.20 * min_dim, high=0.25 * min_dim)
        ndim = len(img_shape)
        indices = [numpy.arange(img_shape[i], dtype=float) * dx[i] for i in range(ndim)]
        if self.randomize_center:
            center = []
            for i in range(ndim):
                while True:
                    center_coord = self.random_state.choice(indices[i])
                    if (center_coord - radius > indices[i][0] and
                            center_coord + radius <= indices[i][-1]):
                        center.append(center_coord)
                        break
            center = numpy.array(center)
        else:
            center = 0.5 * numpy.array(img_shape, dtype=float)
        self.random_state.set_state(state)
        return center, radius

    def initialize(self, img, dx, seed):
        center, radius = self._draw(img.shape, dx, img.ravel()[0])
        indices = numpy.indices(img.shape, dtype=float)
        shape = dx.shape + tuple(numpy.ones(img.ndim, dtype=int))
        indices *= dx.reshape(shape)
        indices -= center.reshape(shape)
        indices **= 2
        return indices.sum(axis=0)**0.5 <= radius


class ThresholdBallInitializer(InitializerBase):
    """ A ball around the seed that just fits into the bright component the seed lies in
    (reference ball.py:118-159).  "Bright" means brighter than the image smoothed with a Gaussian
    of width ``sigma``; if the seed is in no component the nearest component voxel takes its
    place.  The radius is the seed's distance to the component's complement, taken from the device
    marcher where the reference calls ``skfmm.distance``.  Host side, once per image. """

    def __init__(self, sigma=4.0):
        self.sigma = sigma

    @staticmethod
    def _seed_to_index(seed):
        return tuple(numpy.asarray(seed).round().astype(int))

    def initialize(self, img, dx, seed):
        from scipy import ndimage
        from ...util.distance_transform import distance_transform

        components, _ = ndimage.label(img > ndimage.gaussian_filter(img, self.sigma))
        centre = numpy.asarray(seed, dtype=float)
        if components[self._seed_to_index(centre)] == 0:
            voxels = numpy.argwhere(components > 0) * numpy.asarray(dx, dtype=float)
            centre = voxels[numpy.linalg.norm(voxels - centre, axis=1).argmin()]
        at = self._seed_to_index(centre)
        component = components == components[at]
        # the bool component as a level set (0 outside, 1 inside), full distance map
        to_boundary, _ = distance_transform(component.astype(float), band=0, dx=dx)
        grid = numpy.indices(img.shape, dtype=float)
        offsets = [(grid[a] - centre[a]) * dx[a] for a in range(img.ndim)]
        return numpy.sqrt(sum(o * o for o in offsets)) < to_boundary[at]
