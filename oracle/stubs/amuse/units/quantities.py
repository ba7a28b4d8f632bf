from _q import to_quantity
