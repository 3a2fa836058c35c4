#!/bin/bash
# Builds host/wire.c with AddressSanitizer + UBSan and feeds it mutated copies of the fixture messages (tests/golden/wire_vectors.json):
# unpack, pack again, free. usage: tools/fuzz_wire.sh [number of inputs, default 200000]
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
N=${1:-200000}
TMP=$(mktemp -d)
python - "$ROOT" "$TMP/in.bin" "$N" <<'PY'
import json, struct, sys
import numpy as np
root, out, n = sys.argv[1], sys.argv[2], int(sys.argv[3])
vectors = json.load(open(root + "/tests/golden/wire_vectors.json"))["vectors"]
types = {"RxRequest": 0, "TxRequest": 1, "Response": 2, "TxData": 3}
rng = np.random.default_rng(11)
with open(out, "wb") as f:
    for trial in range(n):
        v = vectors[trial % len(vectors)]
        data = bytearray(bytes.fromhex(v["hex"]))
        for _ in range(int(rng.integers(1, 5))):
            kind = int(rng.integers(0, 5))
            if kind == 0 and data:
                data[int(rng.integers(0, len(data)))] ^= 1 << int(rng.integers(0, 8))
            elif kind == 1 and data:
                del data[int(rng.integers(0, len(data))):]
            elif kind == 2 and data:
                at = int(rng.integers(0, len(data)))
                data[at:at] = data[at:at + int(rng.integers(1, 9))]
            elif kind == 3:
                at = int(rng.integers(0, len(data) + 1))
                data[at:at] = bytes(rng.integers(0, 256, int(rng.integers(1, 6)), dtype=np.uint8))
            else:
                data += bytes(rng.integers(0, 256, int(rng.integers(1, 12)), dtype=np.uint8))
        f.write(struct.pack("<BI", types[v["type"]], len(data)) + bytes(data))
PY
gcc -std=gnu99 -g -O1 -fsanitize=address,undefined -fno-omit-frame-pointer -o "$TMP/fuzz" "$ROOT/tools/fuzz_wire.c" "$ROOT/sdr-modem_b200/host/wire.c"
"$TMP/fuzz" "$TMP/in.bin"
rm -rf "$TMP"
