#!/bin/bash
# Measurement aid (GPU box): runs a command once per variant library built by tools/ab_build.sh.
# usage: tools/ab_run.sh "<command>" <name> [<name>...]
cmd=$1; shift
cp sdr-modem_b200/libsdrmodem_b200.so /tmp/lib_orig.so
for name in "$@"; do
    cp tools/ab/$name.so sdr-modem_b200/libsdrmodem_b200.so
    echo "== $name"
    eval "$cmd"
done
cp /tmp/lib_orig.so sdr-modem_b200/libsdrmodem_b200.so
