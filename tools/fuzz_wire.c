/* Measurement aid: AddressSanitizer / UBSan run of the wire codec over mutated messages (tools/fuzz_wire.sh builds and runs it). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include "../include/sdrm/api.h"
/* input file: records of [u8 type][u32 len][bytes] */
int main(int argc, char **argv) {
    FILE *f = fopen(argv[1], "rb");
    size_t n = 0, parsed = 0;
    for (;;) {
        uint8_t type; uint32_t len;
        if (fread(&type, 1, 1, f) != 1 || fread(&len, 4, 1, f) != 1) break;
        uint8_t *buf = malloc(len ? len : 1);
        if (len && fread(buf, 1, len, f) != len) break;
        n++;
        if (type == 0) { RxRequest *m = rx_request__unpack(NULL, len, buf); if (m) { parsed++; size_t s = rx_request__get_packed_size(m); uint8_t *o = malloc(s ? s : 1); rx_request__pack(m, o); free(o); rx_request__free_unpacked(m, NULL);} }
        else if (type == 1) { TxRequest *m = tx_request__unpack(NULL, len, buf); if (m) { parsed++; size_t s = tx_request__get_packed_size(m); uint8_t *o = malloc(s ? s : 1); tx_request__pack(m, o); free(o); tx_request__free_unpacked(m, NULL);} }
        else if (type == 2) { Response *m = response__unpack(NULL, len, buf); if (m) { parsed++; size_t s = response__get_packed_size(m); uint8_t *o = malloc(s ? s : 1); response__pack(m, o); free(o); response__free_unpacked(m, NULL);} }
        else { TxData *m = tx_data__unpack(NULL, len, buf); if (m) { parsed++; size_t s = tx_data__get_packed_size(m); uint8_t *o = malloc(s ? s : 1); tx_data__pack(m, o); free(o); tx_data__free_unpacked(m, NULL);} }
        free(buf);
    }
    printf("%zu inputs, %zu parsed\n", n, parsed);
    return 0;
}
