// Drop-in for zmqpublisher.h (zmqpublisher.h:7-28, .cpp:15-96): same methods, same public
// `connected`, same three-frame wire format and socket options (sdrb_publisher_*).
#ifndef ZMQPUBLISHER_H
#define ZMQPUBLISHER_H
#include <cstdint>
#include <string>

struct sdrb_publisher;

class ZmqPublisher {
public:
    ZmqPublisher();
    ~ZmqPublisher();
    void connect();
    void setAddress(std::string address);
    void setBind(bool b = false);
    void publish(unsigned char *buf, uint32_t len, std::string topic, uint32_t sampleRate);
    bool connected;

private:
    sdrb_publisher *pub;
    std::string bindAddress;
    bool bind;
};
#endif
