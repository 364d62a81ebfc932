// Header-only stand-in for the handful of Qt5 types the SDRReceiver hot-path
// sources touch. TEST INFRASTRUCTURE ONLY: it exists so that oracle/Makefile can
// compile the *unmodified* reference translation units from /root/reference
// (vfo.cpp, halfbanddecimator.cpp, oscillator.cpp, jonti/dsp.cpp, jonti/sdr.cpp,
// gnuradio/firfilter.cpp, zmqpublisher.cpp, sdrj.cpp) without Qt installed.
// Nothing in the product (sdrreceiver_b200/) includes this file.
#ifndef QSHIM_CORE_H
#define QSHIM_CORE_H

#include <cassert>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
#include <sys/types.h>

typedef long long qint64;
typedef unsigned char quint8;

#define Q_OBJECT
#define signals public
#define slots
#define emit
#define SIGNAL(x) #x
#define SLOT(x) #x
#define QT_BEGIN_NAMESPACE
#define QT_END_NAMESPACE

template <class T>
class QVector : public std::vector<T> {
public:
    QVector() {}
    explicit QVector(int n) : std::vector<T>(n) {}
    QVector(int n, const T &v) : std::vector<T>(n, v) {}
    QVector(std::initializer_list<T> il) : std::vector<T>(il) {}
    int length() const { return (int)std::vector<T>::size(); }
    int size() const { return (int)std::vector<T>::size(); }
    int count() const { return size(); }
    const T &at(int i) const { return std::vector<T>::at(i); }
    void append(const T &v) { this->push_back(v); }
    QVector<T> &operator<<(const T &v) { this->push_back(v); return *this; }
    bool contains(const T &v) const {
        for (const T &x : *this) if (x == v) return true;
        return false;
    }
    QVector<T> mid(int pos, int len = -1) const {
        QVector<T> r;
        int n = size();
        if (pos < 0) pos = 0;
        if (len < 0 || pos + len > n) len = n - pos;
        for (int i = 0; i < len; i++) r.push_back((*this)[pos + i]);
        return r;
    }
    static QVector<T> fromStdVector(const std::vector<T> &v) {
        QVector<T> r;
        r.assign(v.begin(), v.end());
        return r;
    }
    std::vector<T> toStdVector() const { return std::vector<T>(this->begin(), this->end()); }
};

template <class T>
class QList : public QVector<T> {
public:
    QList() {}
    QList(std::initializer_list<T> il) : QVector<T>(il) {}
};

class QByteArray {
    std::vector<char> d;
public:
    QByteArray() {}
    QByteArray(const char *s, int n) : d(s, s + n) {}
    void resize(int n) { d.resize(n); }
    int size() const { return (int)d.size(); }
    char &operator[](int i) { return d[i]; }
    const char &operator[](int i) const { return d[i]; }
    char *data() { return d.data(); }
    const char *data() const { return d.data(); }
    // QByteArray::constData() is NUL-terminated in Qt; keep that property.
    const char *constData() const {
        z.assign(d.begin(), d.end());
        return z.c_str();
    }
private:
    mutable std::string z;
};

class QString;
class QStringList;

class QString {
    std::string s;
public:
    QString() {}
    QString(const char *c) : s(c ? c : "") {}
    QString(const std::string &c) : s(c) {}
    int length() const { return (int)s.size(); }
    int size() const { return (int)s.size(); }
    bool isEmpty() const { return s.empty(); }
    QByteArray toUtf8() const { return QByteArray(s.data(), (int)s.size()); }
    std::string toStdString() const { return s; }
    int toInt() const { return (int)strtol(s.c_str(), 0, 10); }
    float toFloat() const { return strtof(s.c_str(), 0); }
    int compare(const QString &o) const { return s.compare(o.s); }
    bool contains(const QString &o) const { return s.find(o.s) != std::string::npos; }
    void chop(int n) { if (n >= (int)s.size()) s.clear(); else s.resize(s.size() - n); }
    static QString number(long long v) { return QString(std::to_string(v)); }
    static QString number(int v) { return QString(std::to_string(v)); }
    static QString number(double v) { std::ostringstream o; o << v; return QString(o.str()); }
    static QString fromLocal8Bit(const char *c) { return QString(c); }
    QString operator+(const QString &o) const { return QString(s + o.s); }
    QString &operator+=(const QString &o) { s += o.s; return *this; }
    bool operator==(const QString &o) const { return s == o.s; }
    bool operator!=(const QString &o) const { return s != o.s; }
    bool operator==(const char *o) const { return s == o; }
    bool operator!=(const char *o) const { return s != o; }
    inline QStringList split(const QString &sep) const;
};
inline QString operator+(const char *a, const QString &b) { return QString(a) + b; }

class QStringList : public QList<QString> {
public:
    QStringList &operator<<(const QString &v) { this->push_back(v); return *this; }
};

inline QStringList QString::split(const QString &sep) const {
    QStringList r;
    size_t p = 0;
    for (;;) {
        size_t q = s.find(sep.s, p);
        if (q == std::string::npos || sep.s.empty()) { r << QString(s.substr(p)); break; }
        r << QString(s.substr(p, q - p));
        p = q + sep.s.size();
    }
    return r;
}

struct QDebugSink {
    template <class T> QDebugSink &operator<<(const T &) { return *this; }
};
inline QDebugSink qDebug() { return QDebugSink(); }

namespace Qt { enum ConnectionType { AutoConnection, DirectConnection, QueuedConnection, UniqueConnection }; }

class QObject {
public:
    QObject(QObject * = 0) {}
    virtual ~QObject() {}
    template <class... A> static bool connect(A...) { return true; }
    template <class... A> bool disconnect(A...) { return true; }
    void deleteLater() {}
};

class QMessageBox {
public:
    void setText(const QString &) {}
    int exec() { return 0; }
};

template <class T>
class QFuture {
public:
    QFuture() {}
    bool isFinished() const { return true; }
    void waitForFinished() {}
    T result() const { return T(); }
};

struct QFutureAny {
    template <class T> operator QFuture<T>() const { return QFuture<T>(); }
};

namespace QtConcurrent {
template <class... A> QFutureAny run(A...) { return QFutureAny(); }
}

class QMutex {
public:
    void lock() {}
    void unlock() {}
};

// The oracle is single-threaded: where the reference's dispatcher thread would go to sleep (jonti/sdr.cpp:152-157:
// buffers_used == 0), wait() throws instead, so a harness that plays both threads gets control back at exactly that point.
struct QShimWouldBlock {};
class QWaitCondition {
public:
    bool wait(QMutex *) { throw QShimWouldBlock(); }
    void wakeAll() {}
};

// In-memory socket: the harness plays rtl_tcp. Bytes it appends to `rx` are what bytesAvailable()/read()/readAll() see,
// what the reference write()s collects in `tx`. qshim_tcp_connect_ok decides what waitForConnected() answers and
// qshim_last_socket points at the socket most recently constructed (sdrj::start_tcp_rtl creates its own, sdrj.cpp:35).
class QTcpSocket;
// one instance across translation units (function-local statics of inline functions are shared)
inline QTcpSocket *&qshim_last_socket_ref() { static QTcpSocket *p = 0; return p; }
inline bool &qshim_tcp_connect_ok_ref() { static bool v = false; return v; }
#define qshim_last_socket (qshim_last_socket_ref())
#define qshim_tcp_connect_ok (qshim_tcp_connect_ok_ref())
class QTcpSocket : public QObject {
public:
    std::string rx, tx, host;
    int port;
    QTcpSocket(QObject * = 0) : port(0) { qshim_last_socket = this; }
    void connectToHost(const QString &h, int p) { host = h.toStdString(); port = p; }
    bool waitForConnected(int) { return qshim_tcp_connect_ok; }
    QString errorString() const { return QString("no network in the oracle"); }
    qint64 bytesAvailable() const { return (qint64)rx.size(); }
    QByteArray readAll() { QByteArray r(rx.data(), (int)rx.size()); rx.clear(); return r; }
    QByteArray read(qint64 n) {
        if (n > (qint64)rx.size()) n = (qint64)rx.size();
        QByteArray r(rx.data(), (int)n);
        rx.erase(0, (size_t)n);
        return r;
    }
    qint64 write(const QByteArray &b) { tx.append(b.data(), (size_t)b.size()); return b.size(); }
    void disconnectFromHost() {}
};

#endif
