/*
 * The streaming hand-off around the DSP chain: the bounded block queue (reference src/queue.c) and the per-session
 * dsp_worker thread (reference src/dsp_worker.c) that drains it through doppler_process_rx and fsk_demod_process.
 */
#define _POSIX_C_SOURCE 200809L

#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "../../include/sdrm/api.h"
#include "../../include/sdrm/doppler.h"
#include "../../include/sdrm/dsp_worker.h"
#include "../../include/sdrm/fsk_demod.h"
#include "../../include/sdrm/queue.h"
#include "sdrm_internal.h"

/* ---- queue ---------------------------------------------------------------------------------------------------------------- */

/* A ring of `size` pinned buffers. Filled blocks occupy [head, head + filled); the block handed to the consumer
 * ("detached", queue.c:168-200) is the slot just before head and returns to the free pool on completion. */
struct queue_t {
    float complex **slots;
    size_t *lens;
    int *pinned;
    uint16_t size;
    uint16_t head;
    uint16_t filled;
    int detached;
    pthread_mutex_t mutex;
    pthread_cond_t condition;
    int poison_pill;
    uint32_t buffer_size;
    bool blocking;
};

int create_queue(uint32_t buffer_size, uint16_t queue_size, bool blocking, queue **out) {
    if (queue_size == 0) {
        SDRM_LOG_ERROR("invalid queue size: %d", queue_size);
        return -1;
    }
    if (buffer_size == 0) {
        SDRM_LOG_ERROR("invalid buffer size: %u", buffer_size);
        return -1;
    }
    struct queue_t *q = calloc(1, sizeof(*q));
    if (q == NULL) {
        return -ENOMEM;
    }
    q->slots = calloc(queue_size, sizeof(float complex *));
    q->lens = calloc(queue_size, sizeof(size_t));
    q->pinned = calloc(queue_size, sizeof(int));
    if (q->slots == NULL || q->lens == NULL || q->pinned == NULL) {
        free(q->slots);
        free(q->lens);
        free(q->pinned);
        free(q);
        return -ENOMEM;
    }
    q->size = queue_size;
    pthread_mutex_init(&q->mutex, NULL);
    pthread_cond_init(&q->condition, NULL);
    q->buffer_size = buffer_size;
    q->blocking = blocking;
    for (uint16_t i = 0; i < queue_size; i++) {
        void *p = NULL;
        /* pinned when a CUDA device is there to DMA from it; ordinary memory otherwise (only the copy speed differs) */
        if (cudaHostAlloc(&p, sizeof(float complex) * buffer_size, cudaHostAllocDefault) == cudaSuccess) {
            q->pinned[i] = 1;
        } else {
            (void) cudaGetLastError();
            p = malloc(sizeof(float complex) * buffer_size);
        }
        if (p == NULL) {
            destroy_queue(q);
            return -ENOMEM;
        }
        q->slots[i] = p;
    }
    *out = q;
    return 0;
}

static uint16_t free_slots(const struct queue_t *q) { return (uint16_t) (q->size - q->filled - (q->detached ? 1 : 0)); }

int queue_put(const float complex *buffer, size_t len, queue *q) {
    if (buffer == NULL || len == 0) {
        return -1;
    }
    if (len > q->buffer_size) {
        SDRM_LOG_ERROR("requested buffer %zu is more than max: %d", len, q->buffer_size);
        return -1;
    }
    pthread_mutex_lock(&q->mutex);
    if (q->blocking) {
        if (q->poison_pill == 1) {
            pthread_mutex_unlock(&q->mutex);
            return -1;
        }
        while (free_slots(q) == 0) {
            pthread_cond_wait(&q->condition, &q->mutex);
            if (q->poison_pill == 1) {
                pthread_mutex_unlock(&q->mutex);
                return -1;
            }
        }
    }
    uint16_t slot;
    if (free_slots(q) == 0) {
        /* full: the newest block is overwritten (queue.c:124-128) */
        SDRM_LOG_ERROR("queue is full");
        if (q->filled == 0) {
            pthread_mutex_unlock(&q->mutex);
            return -1;
        }
        slot = (uint16_t) ((q->head + q->filled - 1) % q->size);
    } else {
        slot = (uint16_t) ((q->head + q->filled) % q->size);
        q->filled++;
    }
    memcpy(q->slots[slot], buffer, sizeof(float complex) * len);
    q->lens[slot] = len;
    pthread_cond_broadcast(&q->condition);
    pthread_mutex_unlock(&q->mutex);
    return 0;
}

void take_buffer_for_processing(float complex **buffer, size_t *len, queue *q) {
    pthread_mutex_lock(&q->mutex);
    while (q->filled == 0) {
        if (q->poison_pill == 1) {
            pthread_mutex_unlock(&q->mutex);
            *buffer = NULL;
            return;
        }
        pthread_cond_wait(&q->condition, &q->mutex);
    }
    *buffer = q->slots[q->head];
    *len = q->lens[q->head];
    q->head = (uint16_t) ((q->head + 1) % q->size);
    q->filled--;
    q->detached = 1;
    pthread_mutex_unlock(&q->mutex);
}

/* internal (rx_group.c): true when take_buffer_for_processing would not block */
int sdrm_queue_has_data(queue *q) {
    pthread_mutex_lock(&q->mutex);
    const int result = q->filled > 0;
    pthread_mutex_unlock(&q->mutex);
    return result;
}

void complete_buffer_processing(queue *q) {
    pthread_mutex_lock(&q->mutex);
    q->detached = 0;
    pthread_cond_broadcast(&q->condition);
    pthread_mutex_unlock(&q->mutex);
}

void interrupt_waiting_the_data(queue *q) {
    if (q == NULL) {
        return;
    }
    pthread_mutex_lock(&q->mutex);
    q->poison_pill = 1;
    pthread_cond_broadcast(&q->condition);
    pthread_mutex_unlock(&q->mutex);
}

void destroy_queue(queue *q) {
    if (q == NULL) {
        return;
    }
    for (uint16_t i = 0; i < q->size; i++) {
        if (q->slots[i] == NULL) {
            continue;
        }
        if (q->pinned[i]) {
            cudaFreeHost(q->slots[i]);
        } else {
            free(q->slots[i]);
        }
    }
    pthread_mutex_destroy(&q->mutex);
    pthread_cond_destroy(&q->condition);
    free(q->slots);
    free(q->lens);
    free(q->pinned);
    free(q);
}

/* ---- dsp_worker -------------------------------------------------------------------------------------------------------------- */

struct dsp_worker_t {
    uint32_t id;
    int client_socket;
    fsk_demod *fsk_demod;
    doppler *dopp;
    queue *queue;
    pthread_t dsp_thread;
    int thread_started;
    FILE *rx_dump_file;
    FILE *demod_file;
    int demod_destination;
};

bool dsp_worker_find_by_id(void *id, void *data) {
    const uint32_t wanted = *(uint32_t *) id;
    return ((dsp_worker *) data)->id == wanted;
}

void dsp_worker_put(float complex *output, size_t output_len, dsp_worker *worker) { queue_put(output, output_len, worker->queue); }

void dsp_worker_shutdown(void *arg, void *data) {
    (void) arg;
    interrupt_waiting_the_data(((dsp_worker *) data)->queue);
}

static int write_all(const uint8_t *buffer, size_t len, int fd) {
    size_t left = len;
    while (left > 0) {
        const ssize_t written = write(fd, buffer + (len - left), left);
        if (written < 0) {
            return -1;
        }
        left -= (size_t) written;
    }
    return 0;
}

/* the loop of reference src/dsp_worker.c:44-106 */
static void *dsp_worker_callback(void *arg) {
    dsp_worker *worker = arg;
    const uint32_t id = worker->id;
    fprintf(stdout, "[%d] dsp_worker is starting\n", id);
    float complex *input = NULL;
    size_t input_len = 0;
    while (true) {
        take_buffer_for_processing(&input, &input_len, worker->queue);
        if (input == NULL) {
            break; /* poison pill */
        }
        if (worker->rx_dump_file != NULL) {
            if (fwrite(input, sizeof(float complex), input_len, worker->rx_dump_file) < input_len) {
                complete_buffer_processing(worker->queue);
                SDRM_LOG_ERROR("[%d] unable to write sdr data", id);
                break;
            }
        }
        if (worker->dopp != NULL) {
            float complex *corrected = NULL;
            size_t corrected_len = 0;
            doppler_process_rx(input, input_len, &corrected, &corrected_len, worker->dopp);
            input = corrected;
            input_len = corrected_len;
        }
        int8_t *symbols = NULL;
        size_t symbols_len = 0;
        if (worker->fsk_demod != NULL && input != NULL) {
            fsk_demod_process(input, input_len, &symbols, &symbols_len, worker->fsk_demod);
        }
        if (symbols == NULL) {
            complete_buffer_processing(worker->queue);
            continue;
        }
        if (worker->demod_file != NULL) {
            if (fwrite(symbols, sizeof(int8_t), symbols_len, worker->demod_file) < symbols_len) {
                complete_buffer_processing(worker->queue);
                SDRM_LOG_ERROR("[%d] unable to write demod data", id);
                break;
            }
        }
        int code = 0;
        if (worker->demod_destination == SDRM_DEMOD_DESTINATION_SOCKET || worker->demod_destination == SDRM_DEMOD_DESTINATION_BOTH) {
            code = write_all((const uint8_t *) symbols, symbols_len, worker->client_socket);
        }
        complete_buffer_processing(worker->queue);
        if (code != 0) {
            break;
        }
    }
    printf("[%d] dsp_worker stopped\n", worker->id);
    return NULL;
}

/* The reference's signature (src/dsp_worker.c:108-197): the fields it reads from the request and the server configuration,
 * handed to sdrm_dsp_worker_create. */
int dsp_worker_create(uint32_t id, int client_socket, struct server_config *server_config, struct RxRequest *req, dsp_worker **worker) {
    if (server_config == NULL || req == NULL || worker == NULL) {
        return -1;
    }
    sdrm_dsp_worker_config c;
    memset(&c, 0, sizeof(c));
    c.rx_center_freq = req->rx_center_freq;
    c.rx_sampling_freq = req->rx_sampling_freq;
    c.rx_dump_file = req->rx_dump_file != 0;
    c.demod_gmsk = req->demod_type == MODEM_TYPE__GMSK;
    c.demod_baud_rate = req->demod_baud_rate;
    c.demod_decimation = req->demod_decimation;
    if (c.demod_gmsk) {
        if (req->fsk_settings == NULL) {
            SDRM_LOG_ERROR("[%d] missing fsk settings", id);
            return -1;
        }
        c.demod_fsk_deviation = req->fsk_settings->demod_fsk_deviation;
        c.demod_fsk_transition_width = req->fsk_settings->demod_fsk_transition_width;
        c.demod_fsk_use_dc_block = req->fsk_settings->demod_fsk_use_dc_block != 0;
    }
    c.demod_destination = (int) req->demod_destination;
    if (req->doppler != NULL) {
        if (req->doppler->n_tle < 3) {
            SDRM_LOG_ERROR("[%d] unable to create doppler correction block", id);
            return -1;
        }
        c.has_doppler = true;
        api_utils_convert_tle(req->doppler->tle, c.doppler_tle);
        /* the scaled integers travel as uint32 on the wire and were written from signed values (api.proto: "degrees times
         * 10^6"); the reference divides the unsigned value, so does this */
        c.doppler_latitude = (int32_t) req->doppler->latitude;
        c.doppler_longitude = (int32_t) req->doppler->longitude;
        c.doppler_altitude = (int32_t) req->doppler->altitude;
        c.doppler_scaled_unsigned = true;
    }
    c.file_start_time_seconds = req->file_settings != NULL ? (int64_t) req->file_settings->start_time_seconds : 0;
    c.buffer_size = server_config->buffer_size;
    c.queue_size = server_config->queue_size;
    c.blocking_queue = server_config->rx_sdr_type == RX_SDR_TYPE_FILE;
    c.base_path = server_config->base_path;
    return sdrm_dsp_worker_create(id, client_socket, &c, worker);
}

int sdrm_dsp_worker_create(uint32_t id, int client_socket, const sdrm_dsp_worker_config *config, dsp_worker **out) {
    if (config == NULL || out == NULL) {
        return -1;
    }
    struct dsp_worker_t *result = calloc(1, sizeof(*result));
    if (result == NULL) {
        return -ENOMEM;
    }
    result->id = id;
    result->client_socket = client_socket;
    int code = 0;
    if (config->has_doppler) {
        char tle[3][80];
        memcpy(tle, config->doppler_tle, sizeof(tle));
        /* same scalings as src/dsp_worker.c:130 */
        double latitude = config->doppler_latitude / 10E6;
        double longitude = config->doppler_longitude / 10E6;
        double altitude = config->doppler_altitude / 10E3;
        if (config->doppler_scaled_unsigned) {
            /* the wire fields are uint32 and the reference divides them as such (src/dsp_worker.c:130) */
            latitude = (uint32_t) config->doppler_latitude / 10E6;
            longitude = (uint32_t) config->doppler_longitude / 10E6;
            altitude = (uint32_t) config->doppler_altitude / 10E3;
        }
        code = doppler_create(latitude, longitude, altitude, config->rx_sampling_freq, config->rx_center_freq, 0, (time_t) config->file_start_time_seconds,
                              config->buffer_size, tle, &result->dopp);
        if (code != 0) {
            SDRM_LOG_ERROR("[%d] unable to create doppler correction block", id);
            dsp_worker_destroy(result);
            return code;
        }
    }
    if (config->demod_gmsk) {
        code = fsk_demod_create(config->rx_sampling_freq, config->demod_baud_rate, config->demod_fsk_deviation,
                                (uint8_t) config->demod_decimation, config->demod_fsk_transition_width,
                                config->demod_fsk_use_dc_block, config->buffer_size, &result->fsk_demod);
    }
    if (code != 0) {
        SDRM_LOG_ERROR("[%d] unable to create demodulator", id);
        dsp_worker_destroy(result);
        return code;
    }
    const char *base = config->base_path != NULL ? config->base_path : ".";
    if (config->rx_dump_file) {
        char path[4096];
        snprintf(path, sizeof(path), "%s/rx.sdr2demod.%d.cf32", base, id);
        result->rx_dump_file = fopen(path, "wb");
        if (result->rx_dump_file == NULL) {
            SDRM_LOG_ERROR("[%d] unable to open file for sdr input: %s", id, path);
            dsp_worker_destroy(result);
            return -1;
        }
    }
    result->demod_destination = config->demod_destination;
    if (config->demod_destination == SDRM_DEMOD_DESTINATION_FILE || config->demod_destination == SDRM_DEMOD_DESTINATION_BOTH) {
        char path[4096];
        snprintf(path, sizeof(path), "%s/rx.demod2client.%d.s8", base, id);
        result->demod_file = fopen(path, "wb");
        if (result->demod_file == NULL) {
            SDRM_LOG_ERROR("[%d] unable to open file for demod output: %s", id, path);
            dsp_worker_destroy(result);
            return -1;
        }
    }
    code = create_queue(config->buffer_size, config->queue_size, config->blocking_queue, &result->queue);
    if (code != 0) {
        dsp_worker_destroy(result);
        return code;
    }
    if (pthread_create(&result->dsp_thread, NULL, &dsp_worker_callback, result) != 0) {
        dsp_worker_destroy(result);
        return -1;
    }
    result->thread_started = 1;
    *out = result;
    return 0;
}

void dsp_worker_destroy(void *data) {
    if (data == NULL) {
        return;
    }
    dsp_worker *worker = data;
    fprintf(stdout, "[%d] dsp_worker is stopping\n", worker->id);
    if (worker->queue != NULL) {
        interrupt_waiting_the_data(worker->queue);
    }
    if (worker->thread_started) {
        pthread_join(worker->dsp_thread, NULL);
    }
    if (worker->queue != NULL) {
        destroy_queue(worker->queue);
    }
    if (worker->rx_dump_file != NULL) {
        fclose(worker->rx_dump_file);
    }
    if (worker->demod_file != NULL) {
        fclose(worker->demod_file);
    }
    fsk_demod_destroy(worker->fsk_demod);
    doppler_destroy(worker->dopp);
    free(worker);
}
