/* Drop-in for reference src/queue.h:10-19: the bounded buffer queue between the SDR thread and a dsp_worker.
 * Same semantics (copy on put, overwrite-last when full unless blocking, detached buffer while it is processed, poison
 * pill), but the buffers are pinned host memory so that the worker's host->device copy is a plain DMA. */
#ifndef SDRM_QUEUE_H
#define SDRM_QUEUE_H

#include <complex.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

typedef struct queue_t queue;

int create_queue(uint32_t buffer_size, uint16_t queue_size, bool blocking, queue **queue);

int queue_put(const float complex *buffer, size_t len, queue *queue);
void take_buffer_for_processing(float complex **buffer, size_t *len, queue *queue);
void complete_buffer_processing(queue *queue);

void interrupt_waiting_the_data(queue *queue);
void destroy_queue(queue *queue);

#endif
