"""CPU-only, build container only (needs /root/reference): the struct layouts this library shares with the reference's callers
are pinned to the reference's own headers — struct server_config (src/server_config.h), the protobuf message structs of the
generated src/api.pb-c.h, struct message_header (src/api.h) and the public struct fir_filter_t (src/dsp/fir_filter.h). Two
programs print sizeof / offsetof of every field, one compiled against the reference's headers, one against include/sdrm/*.h."""
import os
import subprocess
import tempfile

import pytest

from conftest import ROOT

REF = "/root/reference"

FIELDS = {
    "struct server_config": ["bind_address", "port", "read_timeout_seconds", "buffer_size", "queue_size", "rx_sdr_type",
                             "rx_sdr_server_address", "rx_sdr_server_port", "base_path", "rx_file_base_path", "tx_file_base_path",
                             "tx_sdr_type", "tx_plutosdr_gain", "rx_plutosdr_gain", "tx_plutosdr_timeout_millis", "iio"],
    "struct message_header": ["protocol_version", "type", "message_length"],
    "struct DopplerSettings": ["base", "n_tle", "tle", "latitude", "longitude", "altitude"],
    "struct FskDemodulationSettings": ["base", "demod_fsk_deviation", "demod_fsk_transition_width", "demod_fsk_use_dc_block"],
    "struct FskModulationSettings": ["base", "mod_fsk_deviation"],
    "struct FileSettings": ["base", "filename", "start_time_seconds"],
    "struct RxRequest": ["base", "rx_center_freq", "rx_sampling_freq", "rx_dump_file", "rx_offset", "demod_type", "demod_baud_rate",
                         "demod_decimation", "demod_destination", "doppler", "fsk_settings", "file_settings"],
    "struct TxRequest": ["base", "tx_center_freq", "tx_sampling_freq", "tx_dump_file", "tx_offset", "mod_type", "mod_baud_rate",
                         "doppler", "fsk_settings", "file_settings"],
    "struct Response": ["base", "status", "details"],
    "struct TxData": ["base", "data"],
    "struct fir_filter_t": ["decimation", "taps", "aligned_taps_len", "alignment", "taps_len", "original_taps", "working_buffer",
                            "history_offset", "working_len_total", "volk_output", "max_input_buffer_length", "output", "output_len",
                            "num_bytes"],
}


def program(includes):
    lines = ["#include <stddef.h>", "#include <stdio.h>"] + ['#include "%s"' % i for i in includes] + ["int main(void) {"]
    for struct, fields in FIELDS.items():
        lines.append('    printf("%s size %%zu\\n", sizeof(%s));' % (struct, struct))
        for f in fields:
            lines.append('    printf("%s.%s %%zu %%zu\\n", offsetof(%s, %s), sizeof(((%s *) 0)->%s));' % (struct, f, struct, f, struct, f))
    lines += ['    printf("MODEM_TYPE__GMSK %d DEMOD_DESTINATION__BOTH %d RESPONSE_STATUS__FAILURE %d\\n", (int) MODEM_TYPE__GMSK, '
              '(int) DEMOD_DESTINATION__BOTH, (int) RESPONSE_STATUS__FAILURE);',
              '    printf("TYPE_TX_REQUEST %d RESPONSE_DETAILS_RX_IS_BEING_USED %d PROTOCOL_VERSION %d\\n", TYPE_TX_REQUEST, '
              'RESPONSE_DETAILS_RX_IS_BEING_USED, PROTOCOL_VERSION);',
              "    return 0;", "}"]
    return "\n".join(lines) + "\n"


def run(source, include_dirs):
    with tempfile.TemporaryDirectory(dir=os.path.join(ROOT, "tests")) as tmp:
        src = os.path.join(tmp, "layout.c")
        exe = os.path.join(tmp, "layout")
        with open(src, "w") as f:
            f.write(source)
        cmd = ["gcc", "-std=gnu99", "-w", "-o", exe, src] + ["-I" + d for d in include_dirs]
        subprocess.run(cmd, check=True)
        return subprocess.run([exe], check=True, capture_output=True, text=True).stdout


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "src", "server_config.h")), reason="the reference tree is not here")
def test_shared_struct_layouts_equal_the_reference_headers():
    shim = os.path.join(ROOT, "oracle", "shim")
    theirs = run(program(["server_config.h", "api.h", "api.pb-c.h", "dsp/fir_filter.h"]), [shim, os.path.join(REF, "src")])
    ours = run(program(["sdrm/server_config.h", "sdrm/api.h", "sdrm/fir_filter.h"]), [os.path.join(ROOT, "include")])
    assert ours == theirs
    assert "struct RxRequest size 96" in ours and "struct message_header size 6" in ours
