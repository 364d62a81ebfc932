// Declarations of the eight libzmq entry points zmqpublisher.cpp calls, with the
// constant values of libzmq 4.3.x. The oracle harness defines them as an
// in-memory frame capture; tests/test_zmq_wire.py links the real libzmq instead.
#ifndef ORACLE_ZMQ_SHIM_H
#define ORACLE_ZMQ_SHIM_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
#define ZMQ_PUB 1
#define ZMQ_SNDMORE 2
#define ZMQ_RECONNECT_IVL 18
#define ZMQ_RECONNECT_IVL_MAX 21
#define ZMQ_TCP_KEEPALIVE 34
#define ZMQ_TCP_KEEPALIVE_CNT 35
#define ZMQ_TCP_KEEPALIVE_IDLE 36
#define ZMQ_TCP_KEEPALIVE_INTVL 37
void *zmq_ctx_new(void);
void *zmq_socket(void *, int type);
int zmq_setsockopt(void *s, int option, const void *optval, size_t optvallen);
int zmq_bind(void *s, const char *addr);
int zmq_connect(void *s, const char *addr);
int zmq_errno(void);
int zmq_send(void *s, const void *buf, size_t len, int flags);
#ifdef __cplusplus
}
#endif
#endif
