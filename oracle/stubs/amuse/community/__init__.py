from _q import units
