// cognn_shim_net.h -- host message transport between party PROCESSES for the reference-API shim: named, length-prefixed,
// duplex byte-message connections over TCP (loopback by default).  It carries what the reference moves over TaskComm's and
// SCI's host channels (Task-Worker, absent from /root/reference) when the reference's own engine (ss_vertex_centric_algo_kernel.h)
// runs on top of the shim; the device-resident engine does not use it (its plane is NCCL, cognn_b200/host/comm.cpp).
//
//   server side:  Conn* c = net::accept_named(port, "name");     one listening socket per port and process, dispatch by name
//   client side:  Conn* c = net::connect_named(ip, port, "name"); retries until the server listens (parties start in any order)
//   c->send(bytes) queues the message for a sender thread (like osuCrypto::Channel::asyncSend: two parties that send to each other
//   at the same time cannot deadlock on full socket buffers); c->recv(bytes) blocks for one whole message.
#pragma once
#include <arpa/inet.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <sys/socket.h>
#include <unistd.h>

#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace cognn_shim {
namespace net {

[[noreturn]] inline void die(const char* what) {
    printf("cognn_shim net: %s (%s)\n", what, strerror(errno));
    exit(-1);
}
inline void write_all(int fd, const void* p, size_t n) {
    const char* c = (const char*)p;
    while (n) {
        ssize_t k = ::send(fd, c, n, MSG_NOSIGNAL);
        if (k <= 0) {
            if (k < 0 && errno == EINTR) continue;
            die("send failed");
        }
        c += k;
        n -= (size_t)k;
    }
}
inline bool read_all(int fd, void* p, size_t n) {
    char* c = (char*)p;
    while (n) {
        ssize_t k = ::recv(fd, c, n, 0);
        if (k == 0) return false;
        if (k < 0) {
            if (errno == EINTR) continue;
            return false;
        }
        c += k;
        n -= (size_t)k;
    }
    return true;
}

class Conn {
public:
    explicit Conn(int fd) : fd_(fd), sender_([this] { pump(); }) {}
    ~Conn() { close(); }
    void send(std::string msg) {
        {
            std::lock_guard<std::mutex> l(m_);
            q_.push_back(std::move(msg));
        }
        cv_.notify_all();
    }
    void recv(std::string& out) {
        uint64_t n = 0;
        if (!read_all(fd_, &n, 8)) die("peer closed the connection");
        out.resize(n);
        if (n && !read_all(fd_, &out[0], n)) die("peer closed the connection inside a message");
    }
    void close() {
        {
            std::unique_lock<std::mutex> l(m_);
            if (closed_) return;
            cv_.wait(l, [this] { return q_.empty() && !busy_; });  // flush what was queued
            closed_ = true;
        }
        cv_.notify_all();
        if (sender_.joinable()) sender_.join();
        ::shutdown(fd_, SHUT_RDWR);
        ::close(fd_);
    }

private:
    void pump() {
        for (;;) {
            std::string msg;
            {
                std::unique_lock<std::mutex> l(m_);
                cv_.wait(l, [this] { return closed_ || !q_.empty(); });
                if (q_.empty()) return;
                msg = std::move(q_.front());
                q_.pop_front();
                busy_ = true;
            }
            const uint64_t n = msg.size();
            write_all(fd_, &n, 8);
            if (n) write_all(fd_, msg.data(), n);
            {
                std::lock_guard<std::mutex> l(m_);
                busy_ = false;
            }
            cv_.notify_all();
        }
    }
    int fd_;
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<std::string> q_;
    bool closed_ = false, busy_ = false;
    std::thread sender_;
};

// one acceptor per port: accepts, reads the connection's name, parks the socket until somebody asks for that name
class Acceptor {
public:
    explicit Acceptor(int port) {
        lfd_ = ::socket(AF_INET, SOCK_STREAM, 0);
        if (lfd_ < 0) die("socket");
        int one = 1;
        setsockopt(lfd_, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
        sockaddr_in a;
        memset(&a, 0, sizeof(a));
        a.sin_family = AF_INET;
        a.sin_addr.s_addr = htonl(INADDR_ANY);
        a.sin_port = htons((uint16_t)port);
        if (::bind(lfd_, (sockaddr*)&a, sizeof(a)) != 0) die("bind (is another run using the port range?)");
        if (::listen(lfd_, 64) != 0) die("listen");
        th_ = std::thread([this] { loop(); });
        th_.detach();
    }
    int take(const std::string& name) {
        std::unique_lock<std::mutex> l(m_);
        cv_.wait(l, [&] { return parked_.count(name) != 0; });
        int fd = parked_[name];
        parked_.erase(name);
        return fd;
    }

private:
    void loop() {
        for (;;) {
            int fd = ::accept(lfd_, nullptr, nullptr);
            if (fd < 0) {
                if (errno == EINTR) continue;
                return;
            }
            int one = 1;
            setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
            uint32_t n = 0;
            if (!read_all(fd, &n, 4) || n > 4096) {
                ::close(fd);
                continue;
            }
            std::string name(n, '\0');
            if (n && !read_all(fd, &name[0], n)) {
                ::close(fd);
                continue;
            }
            {
                std::lock_guard<std::mutex> l(m_);
                parked_[name] = fd;
            }
            cv_.notify_all();
        }
    }
    int lfd_ = -1;
    std::thread th_;
    std::mutex m_;
    std::condition_variable cv_;
    std::map<std::string, int> parked_;
};

inline Acceptor& acceptor(int port) {
    static std::mutex m;
    static std::map<int, std::unique_ptr<Acceptor>> all;
    std::lock_guard<std::mutex> l(m);
    auto& a = all[port];
    if (!a) a.reset(new Acceptor(port));
    return *a;
}

inline Conn* accept_named(int port, const std::string& name) { return new Conn(acceptor(port).take(name)); }

inline Conn* connect_named(const std::string& ip, int port, const std::string& name, int timeout_s = 120) {
    const auto t0 = std::chrono::steady_clock::now();
    for (;;) {
        int fd = ::socket(AF_INET, SOCK_STREAM, 0);
        if (fd < 0) die("socket");
        sockaddr_in a;
        memset(&a, 0, sizeof(a));
        a.sin_family = AF_INET;
        a.sin_port = htons((uint16_t)port);
        if (inet_pton(AF_INET, ip.c_str(), &a.sin_addr) != 1) die("bad ip address");
        if (::connect(fd, (sockaddr*)&a, sizeof(a)) == 0) {
            int one = 1;
            setsockopt(fd, IPPROTO_TCP, TCP_NODELAY, &one, sizeof(one));
            const uint32_t n = (uint32_t)name.size();
            write_all(fd, &n, 4);
            write_all(fd, name.data(), n);
            return new Conn(fd);
        }
        ::close(fd);
        if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(timeout_s)) die("connect timed out");
        std::this_thread::sleep_for(std::chrono::milliseconds(20));
    }
}

// port base of a run (several test runs can share a host): COGNN_SHIM_PORT_BASE, default 21000
inline int port_base() {
    const char* e = getenv("COGNN_SHIM_PORT_BASE");
    return e ? atoi(e) : 21000;
}

}  // namespace net
}  // namespace cognn_shim
