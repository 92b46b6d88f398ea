"""The `loom.query` front end (loom/query.py) over the B200 query server process.

Same names and argument meaning as the reference: `get_server(root, config)` starts `loom_query ROOT - CONFIG - --none`
(loom.runner.query with block=False, loom/runner.py:346-381) and returns a `QueryServer` with `sample`, `score`,
`batch_score`, `entropy`, `mutual_information` and `score_derivative` (loom/query.py:121-289).  Rows are lists with
one entry per feature in loom's feature order -- `bool` for booleans, `int` for counts, `float` for reals, `None` for
unobserved (data_row_to_protobuf, loom/query.py:64-83).  Errors reported by the server raise, as in the reference.

The requests are `Query.Request` messages (src/schema.proto:342-418) written and parsed here directly -- the wire format
of a handful of small messages -- so the package needs no protobuf runtime.  The computation is the server's
(loom_b200/host/query_server.hpp on libloomb200.so); there is no Python fallback: without the binary `get_server` raises.
"""
import os
import shutil
import struct
import subprocess
import tempfile
import uuid
from collections import namedtuple

from . import pbio

DEFAULTS = {                       # loom/query.py:40-45
    'sample_sample_count': 10,
    'entropy_sample_count': 1000,
    'mutual_information_sample_count': 1000,
    'similar_row_limit': 1000,
}
BUFFER_SIZE = 10

Estimate = namedtuple('Estimate', ['mean', 'variance'])

NONE, SPARSE, DENSE, ALL = 0, 1, 2, 3          # ProductValue.Observed.Sparsity (src/schema.proto:76-81)
OTHER = 0xFFFFFFFF                              # a sampled DPD value outside the dictionary (distributions' OTHER())


# ---------------------------------------------------------------------------------------------- wire format
def _varint(v):
    out = bytearray()
    v = int(v)
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _field_varint(number, v):
    return _varint(number << 3) + _varint(v)


def _field_bytes(number, payload):
    return _varint(number << 3 | 2) + _varint(len(payload)) + payload


def _field_float(number, v):
    return _varint(number << 3 | 5) + struct.pack('<f', v)


def _observed(sparsity, dense=(), sparse=()):
    out = _field_varint(1, sparsity)
    for b in dense:
        out += _field_varint(2, 1 if b else 0)
    for i in sparse:
        out += _field_varint(3, i)
    return out


def _value(data_row):
    """ProductValue of a data row: DENSE mask over all features, values packed by type (loom/query.py:64-83)"""
    if all(v is None for v in data_row):
        return _field_bytes(1, _observed(NONE))
    out = _field_bytes(1, _observed(DENSE, dense=[v is not None for v in data_row]))
    booleans = [v for v in data_row if isinstance(v, bool)]
    counts = [v for v in data_row if isinstance(v, int) and not isinstance(v, bool)]
    reals = [v for v in data_row if isinstance(v, float)]
    if len(booleans) + len(counts) + len(reals) != sum(v is not None for v in data_row):
        raise TypeError("row values must be bool, int, float or None")
    for v in booleans:
        out += _field_varint(2, 1 if v else 0)
    for v in counts:
        out += _field_varint(3, v)
    for v in reals:
        out += _field_float(4, v)
    return out


def _diff(data_row):
    return _field_bytes(1, _value(data_row)) + _field_bytes(2, _field_bytes(1, _observed(NONE)))


def _fields(buf):
    """(number, wire type, value) of every field of a message; length-delimited values as bytes"""
    pos, n = 0, len(buf)
    while pos < n:
        key = shift = 0
        while True:
            b = buf[pos]
            pos += 1
            key |= (b & 0x7F) << shift
            shift += 7
            if not b & 0x80:
                break
        number, wt = key >> 3, key & 7
        if wt == 0:
            v = shift = 0
            while True:
                b = buf[pos]
                pos += 1
                v |= (b & 0x7F) << shift
                shift += 7
                if not b & 0x80:
                    break
            yield number, wt, v
        elif wt == 2:
            length = shift = 0
            while True:
                b = buf[pos]
                pos += 1
                length |= (b & 0x7F) << shift
                shift += 7
                if not b & 0x80:
                    break
            yield number, wt, bytes(buf[pos:pos + length])
            pos += length
        elif wt == 5:
            yield number, wt, bytes(buf[pos:pos + 4])
            pos += 4
        elif wt == 1:
            yield number, wt, bytes(buf[pos:pos + 8])
            pos += 8
        else:
            raise ValueError("unsupported wire type %d" % wt)


def _packed_varints(payload):
    out, pos = [], 0
    while pos < len(payload):
        v = shift = 0
        while True:
            b = payload[pos]
            pos += 1
            v |= (b & 0x7F) << shift
            shift += 7
            if not b & 0x80:
                break
        out.append(v)
    return out


def _parse_value(buf, feature_count):
    """ProductValue -> (observed positions, booleans, counts, reals)"""
    sparsity, dense, sparse, booleans, counts, reals = NONE, [], [], [], [], []
    for number, wt, v in _fields(buf):
        if number == 1:
            for n2, wt2, v2 in _fields(v):
                if n2 == 1:
                    sparsity = v2
                elif n2 == 2:
                    dense += [bool(x) for x in ([v2] if wt2 == 0 else _packed_varints(v2))]
                elif n2 == 3:
                    sparse += [v2] if wt2 == 0 else _packed_varints(v2)
        elif number == 2:
            booleans += [bool(x) for x in ([v] if wt == 0 else _packed_varints(v))]
        elif number == 3:
            counts += [v] if wt == 0 else _packed_varints(v)
        elif number == 4:
            reals += [struct.unpack('<f', v)[0]] if wt == 5 else [x[0] for x in struct.iter_unpack('<f', v)]
    if sparsity == ALL:
        positions = list(range(feature_count))
    elif sparsity == DENSE:
        positions = [i for i, b in enumerate(dense) if b]
    elif sparsity == SPARSE:
        positions = sparse
    else:
        positions = []
    return positions, booleans, counts, reals


def _parse_response(buf):
    r = {'id': '', 'error': [], 'samples': [], 'score': None, 'means': [], 'variances': [], 'ids': [], 'score_diffs': []}
    for number, wt, v in _fields(buf):
        if number == 1:
            r['id'] = v.decode()
        elif number == 2:
            r['error'].append(v.decode())
        elif number == 3:
            for n2, _, v2 in _fields(v):
                if n2 == 1:
                    pos = next((v3 for n3, _, v3 in _fields(v2) if n3 == 1), b'')
                    r['samples'].append(pos)
        elif number == 4:
            for n2, _, v2 in _fields(v):
                if n2 == 1:
                    r['score'] = struct.unpack('<f', v2)[0]
        elif number == 5:
            for n2, wt2, v2 in _fields(v):
                vals = [struct.unpack('<f', v2)[0]] if wt2 == 5 else [x[0] for x in struct.iter_unpack('<f', v2)]
                (r['means'] if n2 == 1 else r['variances']).extend(vals)
        elif number == 6:
            for n2, wt2, v2 in _fields(v):
                if n2 == 1:
                    r['ids'] += [v2] if wt2 == 0 else _packed_varints(v2)
                else:
                    r['score_diffs'] += [struct.unpack('<f', v2)[0]] if wt2 == 5 else [x[0] for x in struct.iter_unpack('<f', v2)]
    return r


# ---------------------------------------------------------------------------------------------- the server process
class ProtobufServer(object):
    """loom/query.py:292-326: the piped `loom_query` process"""

    def __init__(self, root, config=None, exe=None, device=None):
        self.root = str(root)
        exe = exe or shutil.which('loom_query') or os.path.join(os.path.dirname(os.path.abspath(__file__)), 'host', 'loom_query')
        if not os.path.exists(exe):
            raise RuntimeError("loom_query not found (make -C loom_b200/host); there is no Python fallback")
        self._tmp = None
        if config is None or isinstance(config, dict):
            self._tmp = tempfile.mkdtemp(prefix='loom_b200_query_')
            path = os.path.join(self._tmp, 'config.pb.gz')
            pbio.config_write(path, pbio.config_from_dict(config))
            config = path
        env = dict(os.environ)
        if device is not None:
            env['LOOM_B200_DEVICE'] = str(device)
        self.proc = subprocess.Popen([exe, self.root, '-', str(config), '-', '--none'], stdin=subprocess.PIPE,
                                     stdout=subprocess.PIPE, env=env)

    def send(self, request):
        self.proc.stdin.write(struct.pack('<I', len(request)) + request)
        self.proc.stdin.flush()

    def receive(self):
        head = self.proc.stdout.read(4)
        if len(head) != 4:
            raise RuntimeError("loom_query ended without answering (exit status %s)" % self.proc.poll())
        (n,) = struct.unpack('<I', head)
        return _parse_response(self.proc.stdout.read(n))

    def close(self):
        if self.proc.stdin:
            self.proc.stdin.close()
        self.proc.wait()
        if self.proc.stdout:
            self.proc.stdout.close()
        if self._tmp:
            shutil.rmtree(self._tmp, ignore_errors=True)
            self._tmp = None

    def __enter__(self):
        return self

    def __exit__(self, *unused):
        self.close()


class QueryServer(object):
    """loom/query.py:101-289"""

    def __init__(self, protobuf_server):
        self.protobuf_server = protobuf_server

    @property
    def root(self):
        return self.protobuf_server.root

    def close(self):
        self.protobuf_server.close()

    def __enter__(self):
        return self

    def __exit__(self, *unused):
        self.close()

    def _call(self, body):
        request = _field_bytes(1, str(uuid.uuid4()).encode()) + body
        self.protobuf_server.send(request)
        response = self.protobuf_server.receive()
        if response['error']:
            raise Exception('\n'.join(response['error']))
        return response

    @staticmethod
    def _data_out(value, feature_count):
        """protobuf_to_data_row (loom/query.py:86-94) for a sampled ProductValue.  loom's feature order is booleans |
        counts | reals and the observed positions ascend, so the packed values concatenated in that order line up."""
        positions, booleans, counts, reals = _parse_value(value, feature_count)
        out = [None] * feature_count
        for i, v in zip(positions, list(booleans) + list(counts) + list(reals)):
            out[i] = v
        return out

    def sample(self, to_sample, conditioning_row=None, sample_count=None):
        if sample_count is None:
            sample_count = DEFAULTS['sample_sample_count']
        if conditioning_row is None:
            conditioning_row = [None for _ in to_sample]
        assert len(to_sample) == len(conditioning_row)
        body = _field_bytes(2, _field_bytes(1, _diff(conditioning_row)) +
                            _field_bytes(2, _observed(DENSE, dense=to_sample)) + _field_varint(3, sample_count))
        response = self._call(body)
        samples = []
        for value in response['samples']:
            data_out = self._data_out(value, len(to_sample))
            for i, val in enumerate(data_out):
                if val is None:
                    assert not to_sample[i]
                    data_out[i] = conditioning_row[i]
            samples.append(data_out)
        return samples

    def _send_score(self, row):
        self.protobuf_server.send(_field_bytes(1, str(uuid.uuid4()).encode()) + _field_bytes(3, _field_bytes(1, _diff(row))))

    def _receive_score(self):
        response = self.protobuf_server.receive()
        if response['error']:
            raise Exception('\n'.join(response['error']))
        return response['score']

    def score(self, row):
        self._send_score(row)
        return self._receive_score()

    def batch_score(self, rows, buffer_size=BUFFER_SIZE):
        buffered = 0
        for row in rows:
            self._send_score(row)
            if buffered < buffer_size:
                buffered += 1
            else:
                yield self._receive_score()
        for _ in range(buffered):
            yield self._receive_score()

    def entropy(self, row_sets, col_sets, conditioning_row=None, sample_count=None):
        row_sets = list(set(map(frozenset, row_sets)) | set([frozenset()]))
        col_sets = list(set(map(frozenset, col_sets)) | set([frozenset()]))
        if sample_count is None:
            sample_count = DEFAULTS['entropy_sample_count']
        body = b''
        for feature_set in row_sets:
            body += _field_bytes(1, _observed(SPARSE, sparse=sorted(feature_set)))
        for feature_set in col_sets:
            body += _field_bytes(2, _observed(SPARSE, sparse=sorted(feature_set)))
        body += _field_bytes(3, _diff(conditioning_row if conditioning_row is not None else []))
        body += _field_varint(4, sample_count)
        response = self._call(_field_bytes(4, body))
        means, variances = response['means'], response['variances']
        size = len(row_sets) * len(col_sets)
        assert len(means) == size, means
        assert len(variances) == size, variances
        means, variances = iter(means), iter(variances)
        return {row_set | col_set: Estimate(next(means), next(variances)) for row_set in row_sets for col_set in col_sets}

    def mutual_information(self, feature_set1, feature_set2, entropys=None, conditioning_row=None, sample_count=None):
        """mutual information between feature_set1 and feature_set2 conditioned on conditioning_row"""
        feature_set1, feature_set2 = frozenset(feature_set1), frozenset(feature_set2)
        if sample_count is None:
            sample_count = DEFAULTS['mutual_information_sample_count']
        feature_union = frozenset.union(feature_set1, feature_set2)
        if entropys is None:
            entropys = self.entropy([feature_set1], [feature_set2], conditioning_row, sample_count)
        mi = entropys[feature_set1].mean + entropys[feature_set2].mean - entropys[feature_union].mean
        variance = entropys[feature_set1].variance + entropys[feature_set2].variance + entropys[feature_union].variance
        return Estimate(mi, variance)

    def score_derivative(self, update_row, score_rows=None, row_limit=None):
        if row_limit is None:
            row_limit = DEFAULTS['similar_row_limit']
        body = b''
        if score_rows is not None:
            for data_row in score_rows:
                body += _field_bytes(1, _diff(data_row))
        body += _field_bytes(2, _diff(update_row)) + _field_varint(3, row_limit)
        response = self._call(_field_bytes(5, body))
        return list(zip(response['ids'], response['score_diffs']))


def get_server(root, config=None, exe=None, device=None):
    """loom/query.py:329-331; `config` is a config.pb[.gz] path, a dict shaped like loom/config.py DEFAULTS, or None"""
    return QueryServer(ProtobufServer(root, config, exe=exe, device=device))
