"""A small proto2 parser and google.protobuf message classes for loom's file formats.

There is no `protoc` in the image, so the test-side message classes are built at run time from
.proto TEXT: tests/golden/loom_wire.proto (our restatement of the messages the native codec
handles) + tests/golden/distributions_standin.proto.  google.protobuf is then an independent
implementation of the wire format to check loom_b200/io/lb_io.cc against.  Test infrastructure.
"""
import os
import re

from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")

SCALARS = {"double": 1, "float": 2, "int64": 3, "uint64": 4, "int32": 5, "fixed64": 6, "fixed32": 7, "bool": 8,
           "string": 9, "bytes": 12, "uint32": 13, "sfixed32": 15, "sfixed64": 16, "sint32": 17, "sint64": 18}
LABELS = {"optional": 1, "required": 2, "repeated": 3}


def tokenize(text):
    text = re.sub(r"//[^\n]*", "", text)
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.findall(r'"[^"]*"|[A-Za-z_][A-Za-z0-9_.]*|-?\d+|[{}\[\];=]', text)


def parse(text):
    """-> {"package": str, "imports": [...], "messages": [msg]}; msg = {"name", "fields": [(label, type, name,
    number)], "messages": [...], "enums": [{"name", "values": [(name, number)]}]}"""
    toks = tokenize(text)
    pos = 0

    def peek():
        return toks[pos] if pos < len(toks) else None

    def take(expected=None):
        nonlocal pos
        t = toks[pos]
        if expected is not None and t != expected:
            raise ValueError("expected %r, got %r at token %d" % (expected, t, pos))
        pos += 1
        return t

    def parse_enum():
        name = take()
        take("{")
        values = []
        while peek() != "}":
            vname = take()
            take("=")
            values.append((vname, int(take())))
            take(";")
        take("}")
        if peek() == ";":
            take()
        return {"name": name, "values": values}

    def parse_message():
        name = take()
        take("{")
        msg = {"name": name, "fields": [], "messages": [], "enums": []}
        while peek() != "}":
            t = take()
            if t == "message":
                msg["messages"].append(parse_message())
            elif t == "enum":
                msg["enums"].append(parse_enum())
            elif t in LABELS:
                ftype, fname = take(), take()
                take("=")
                number = int(take())
                if peek() == "[":
                    while take() != "]":
                        pass
                take(";")
                msg["fields"].append((t, ftype, fname, number))
            elif t == ";":
                pass
            else:
                raise ValueError("unexpected token %r in message %s" % (t, name))
        take("}")
        if peek() == ";":
            take()
        return msg

    out = {"package": "", "imports": [], "messages": []}
    while pos < len(toks):
        t = take()
        if t == "package":
            out["package"] = take()
            take(";")
        elif t == "import":
            out["imports"].append(take().strip('"'))
            take(";")
        elif t == "message":
            out["messages"].append(parse_message())
        elif t == "option":
            while take() != ";":
                pass
        elif t == ";":
            pass
        else:
            raise ValueError("unexpected top-level token %r" % t)
    return out


def flatten(parsed):
    """{full message name: {field name: (label, type, number)}} and {full enum name: {value: number}}"""
    msgs, enums = {}, {}

    def walk(msg, prefix):
        full = prefix + "." + msg["name"] if prefix else msg["name"]
        msgs[full] = {name: (label, ftype, number) for label, ftype, name, number in msg["fields"]}
        for e in msg["enums"]:
            enums[full + "." + e["name"]] = dict(e["values"])
        for m in msg["messages"]:
            walk(m, full)

    for m in parsed["messages"]:
        walk(m, parsed["package"])
    return msgs, enums


def _known_names(parsed_files):
    msgs, enums = set(), set()
    for p in parsed_files:
        m, e = flatten(p)
        msgs.update(m)
        enums.update(e)
    return msgs, enums


def _resolve(name, scope, msgs, enums):
    parts = scope.split(".")
    for depth in range(len(parts), -1, -1):
        cand = ".".join(parts[:depth] + [name])
        if cand in msgs:
            return "." + cand, 11
        if cand in enums:
            return "." + cand, 14
    raise KeyError("cannot resolve type %s in scope %s" % (name, scope))


def to_file_descriptor(parsed, filename, msgs, enums):
    fdp = descriptor_pb2.FileDescriptorProto()
    fdp.name = filename
    fdp.package = parsed["package"]
    fdp.syntax = "proto2"
    for dep in parsed["imports"]:
        fdp.dependency.append(dep)

    def fill(dp, msg, scope):
        dp.name = msg["name"]
        full = scope + "." + msg["name"]
        for label, ftype, fname, number in msg["fields"]:
            f = dp.field.add()
            f.name, f.number, f.label = fname, number, LABELS[label]
            if ftype in SCALARS:
                f.type = SCALARS[ftype]
            else:
                f.type_name, f.type = _resolve(ftype, full, msgs, enums)
        for e in msg["enums"]:
            ep = dp.enum_type.add()
            ep.name = e["name"]
            for vname, number in e["values"]:
                v = ep.value.add()
                v.name, v.number = vname, number
        for m in msg["messages"]:
            fill(dp.nested_type.add(), m, full)

    for m in parsed["messages"]:
        fill(fdp.message_type.add(), m, parsed["package"])
    return fdp


_classes = None


def messages():
    """{"Row": class, "ProductValue": class, ...} for package protobuf.loom, plus "dist.X" for distributions"""
    global _classes
    if _classes is None:
        dist = parse(open(os.path.join(GOLDEN, "distributions_standin.proto")).read())
        loom = parse(open(os.path.join(GOLDEN, "loom_wire.proto")).read())
        msgs, enums = _known_names([dist, loom])
        pool = descriptor_pool.DescriptorPool()
        pool.Add(to_file_descriptor(dist, "distributions_standin.proto", msgs, enums))
        pool.Add(to_file_descriptor(loom, "loom_wire.proto", msgs, enums))
        out = {}
        for full in msgs:
            cls = message_factory.GetMessageClass(pool.FindMessageTypeByName(full))
            if full.startswith("protobuf.loom."):
                out[full[len("protobuf.loom."):]] = cls
            else:
                out["dist." + full[len("protobuf.distributions."):]] = cls
        _classes = out
    return _classes


# ---- loom's file framing, in Python (src/protobuf_stream.hpp): used to write test inputs and read outputs
import gzip
import struct


def _open(path, mode):
    return gzip.open(path, mode) if str(path).endswith(".gz") else open(path, mode)


def dump_message(path, message):
    with _open(path, "wb") as f:
        f.write(message.SerializeToString())


def load_message(path, cls):
    with _open(path, "rb") as f:
        m = cls()
        m.ParseFromString(f.read())
        return m


def dump_stream(path, msgs):
    with _open(path, "wb") as f:
        for m in msgs:
            data = m.SerializeToString()
            f.write(struct.pack("<I", len(data)))
            f.write(data)


def load_stream(path, cls):
    out = []
    with _open(path, "rb") as f:
        while True:
            head = f.read(4)
            if not head:
                return out
            (n,) = struct.unpack("<I", head)
            m = cls()
            m.ParseFromString(f.read(n))
            out.append(m)
