"""Minimal PHP unserialize() reader for the reference's test/test_NNN/model.bin files.

model.bin = PHP serialize() of array(run -> array(query_idx -> result)) written by the
reference's test harness (test/helpers.inc, model load at :3972).  Test tooling only.
"""


def php_unserialize(data: bytes):
    pos = 0

    def read_until(ch):
        nonlocal pos
        end = data.index(ch, pos)
        s = data[pos:end]
        pos = end + 1
        return s

    def parse():
        nonlocal pos
        t = data[pos:pos + 1]
        if t == b'N':
            pos += 2
            return None
        pos += 2  # type + ':'
        if t == b'i':
            return int(read_until(b';'))
        if t == b'd':
            s = read_until(b';').decode()
            return float(s)
        if t == b'b':
            return read_until(b';') == b'1'
        if t == b's':
            n = int(read_until(b':'))
            assert data[pos:pos + 1] == b'"'
            s = data[pos + 1:pos + 1 + n]
            pos += n + 3  # quotes + ';'
            try:
                return s.decode('utf-8')
            except UnicodeDecodeError:
                return s.decode('latin-1')
        if t == b'a':
            n = int(read_until(b':'))
            assert data[pos:pos + 1] == b'{'
            pos += 1
            out = {}
            for _ in range(n):
                k = parse()
                v = parse()
                out[k] = v
            assert data[pos:pos + 1] == b'}'
            pos += 1
            return out
        raise ValueError(f"unknown type {t!r} at {pos}")

    return parse()
