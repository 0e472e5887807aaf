def to_rgba_array(c):
    return [c]


def to_hex(c):
    return "#000000"
