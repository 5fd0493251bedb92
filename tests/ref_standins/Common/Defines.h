// Minimal stand-in for cryptoTools' Common/Defines.h + Network/* (TEST ONLY: a real CoGNN build links libOTe / cryptoTools).
// include/engine.h:157-201 and include/comm_sync.h only need named sessions between party processes with
// Channel::asyncSend(std::string&&) / recv(std::string&) / close(); they run here over cognn_shim_net.h.
#pragma once
#include <list>
#include <memory>
#include <string>

#include "../../../cognn_b200/host/shim/cognn_shim_net.h"

namespace osuCrypto {

class IOService {
public:
    explicit IOService(int = 0) {}
    void stop() {}
};
enum class SessionMode { Client, Server };

class Channel {
public:
    Channel() {}
    explicit Channel(cognn_shim::net::Conn* c) : c_(c) {}
    void asyncSend(std::string&& s) { c_->send(std::move(s)); }
    void asyncSend(const std::string& s) { c_->send(s); }
    void send(const std::string& s) { c_->send(s); }
    void recv(std::string& s) { c_->recv(s); }
    void close() {
        if (c_) c_->close();
    }

private:
    std::shared_ptr<cognn_shim::net::Conn> c_;
};

class Session {
public:
    Session(IOService&, const std::string& ip, uint32_t port, SessionMode mode, const std::string& name)
        : ip_(ip), port_((int)port), mode_(mode), name_(name) {}
    Channel addChannel(const std::string& a = "", const std::string& b = "") {
        const std::string n = name_ + "/" + a + "/" + b;
        // the reference's port range (1712 + tile, engine.h:158) is shifted into the shim's range so that parallel test runs
        // with different COGNN_SHIM_PORT_BASE do not collide
        const int port = cognn_shim::net::port_base() + 100 + (port_ - 1712);
        return Channel(mode_ == SessionMode::Server ? cognn_shim::net::accept_named(port, n)
                                                    : cognn_shim::net::connect_named(ip_, port, n));
    }
    void stop() {}

private:
    std::string ip_;
    int port_;
    SessionMode mode_;
    std::string name_;
};
typedef Session Endpoint;

}  // namespace osuCrypto
