"""Restated `pypose.optim.strategy.TrustRegion` (0.6.8)."""


class TrustRegion(object):
    def __init__(self, radius=1e6, high=.5, low=1e-3, up=2., down=.5, factor=.5, max=1e5, min=1e-3):
        assert radius > 0 and 0 < low < high < 1 and up > 1 and 0 < down < 1 and 0 < factor < 1
        assert 0 < min < max
        self.min, self.max, self.down = min, max, down
        self.defaults = {'radius': radius, 'high': high, 'low': low, 'up': up,
                         'down': down, 'factor': factor, 'damping': 1 / radius}

    def update(self, pg, last, loss, J, D, R, *args, **kwargs):
        quality = (last - loss) / -((J @ D).mT @ (2 * R + J @ D)).squeeze()
        pg['radius'] = 1. / pg['damping']
        if quality > pg['high']:
            pg['radius'] = pg['up'] * pg['radius']
            pg['down'] = self.down
        elif quality > pg['low']:
            pg['radius'] = pg['radius']
            pg['down'] = self.down
        else:
            pg['radius'] = pg['radius'] * pg['down']
            pg['down'] = pg['down'] * pg['factor']
        pg['down'] = max(self.min, min(pg['down'], self.max))
        pg['radius'] = max(self.min, min(pg['radius'], self.max))
        pg['damping'] = 1. / pg['radius']
