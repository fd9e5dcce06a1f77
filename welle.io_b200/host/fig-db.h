/*
 * fig-db.h — the service database of the host glue: reads the Fast Information Groups of CRC-checked FIBs and answers the
 * RadioReceiver getters (getServiceList, getComponents, getSubchannel, getEnsemble*).  Behavioural reference: the reference's
 * FIBProcessor (backend/fib-processor.cpp), whose observable semantics are kept:
 *   - FIG 0/0 ensemble id + change flag (:120-159), FIG 0/1 sub-channel organisation into a 64-entry table (:161-245),
 *     FIG 0/2 services and components (:247-358) with the two-sightings / one-decrement-per-second acceptance rule (:285-327),
 *     first definition of a (SId, component number) wins (:1135-1219), FIG 0/3 packet components (:360-395), FIG 0/5 sub-channel
 *     language (:397-432), FIG 0/9 ECC + local time offset and FIG 0/10 date/time (:485-531), FIG 0/14 FEC scheme (:611-628),
 *     FIG 0/17 programme type / language (:630-664), FIG 1/0, 1/1, 1/4, 1/5 labels (:760-895), FIG 2/0, 2/1, 2/4, 2/5 extended
 *     label segments (:897-1083);
 *   - a FIG of type 7 ends the FIB, unknown types are skipped by their length (:52-78); FIG 1 with the OE flag is ignored (:778).
 * Own code: works on the 30 data bytes of a FIB (the reference walks a one-bit-per-byte array).  Pure host C++, no CUDA, no
 * dependency on the C ABI: tests/test_figdb.py feeds the same FIBs to this class and to the compiled reference and compares
 * the dumps.  Not thread safe by itself: the owner serialises access (FigDb::m).
 */
#ifndef DABB_HOST_FIG_DB_H
#define DABB_HOST_FIG_DB_H
#include "dab_api.h"

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace dabb_host {

/* ETSI EN 300 401 Table 8 (short-form sub-channel sizes): size in CU, protection level (dab-constants.cpp:45-109) */
static const int16_t kUepSizeCu[64] = {16,21,24,29,35, 24,29,35,42,52, 29,35,42,52, 32,42,48,58,70, 40,52,58,70,84, 48,58,70,84,104, 58,70,84,104,
                                       64,84,96,116,140, 80,104,116,140,168, 96,116,140,168,208, 116,140,168,208,232, 128,168,192,232,280, 160,208,280, 192,280,416};
static const int8_t kUepProtLevel[64] = {5,4,3,2,1, 5,4,3,2,1, 5,4,3,2, 5,4,3,2,1, 5,4,3,2,1, 5,4,3,2,1, 5,4,3,2, 5,4,3,2,1, 5,4,3,2,1, 5,4,3,2,1, 5,4,3,2,1, 5,4,3,2,1, 5,4,2, 5,3,1};

struct FigEvents {           /* what one FIB asks the owner to signal, in order of occurrence */
    enum Kind { NewEnsemble, ServiceDetected, EnsembleLabel, RestartService, DateTime };
    struct Ev { Kind kind; uint32_t id; };
    std::vector<Ev> list;
};

class FigDb {
public:
    std::mutex m;
    uint16_t eid = 0; uint8_t ecc = 0; DabLabel ensLabel;
    std::vector<Service> services;                       /* in order of acceptance, like the reference's vector */
    std::vector<ServiceComponent> components;
    std::vector<Subchannel> subch = std::vector<Subchannel>(64);
    dab_date_time_t dateTime; bool timeOffsetReceived = false;
    std::map<uint32_t, int8_t> repeatCount;
    std::chrono::steady_clock::time_point lastDecrement = std::chrono::steady_clock::now();
    std::chrono::system_clock::time_point timeLastFCT0Frame = std::chrono::system_clock::now();     /* fib-processor.cpp:141,1272 */

    void clear()
    {
        std::lock_guard<std::mutex> l(m);
        components.clear(); subch.assign(64, Subchannel()); services.clear(); repeatCount.clear();
        lastDecrement = std::chrono::steady_clock::now();
        timeLastFCT0Frame = std::chrono::system_clock::now();
    }

    Service* findService(uint32_t sid) { for (auto& s : services) if (s.serviceId == sid) return &s; return nullptr; }
    ServiceComponent* findComponent(uint32_t sid, int scids) { for (auto& c : components) if (c.SId == sid && c.componentNr == scids) return &c; return nullptr; }
    ServiceComponent* findPacketComponent(int scid) { for (auto& c : components) if (c.TMid == 3 && c.SCId == scid) return &c; return nullptr; }

    /* one FIB: b = its 32 bytes (30 data bytes + CRC, already checked by the caller).  Like the reference, a field of a malformed FIG
     * that runs past the data bytes reads the CRC bytes; nothing is read beyond the FIB */
    void parseFib(const uint8_t* b, FigEvents& ev)
    {
        std::lock_guard<std::mutex> l(m);
        int p = 0;
        while (p < 30) {
            const int type = b[p] >> 5, len = b[p] & 0x1F;
            if (type == 7) return;
            /* a FIG never extends beyond the FIB; fields are read through rd() which yields zero bits past the end */
            fig_ = b + p; figAvail_ = 32 - p;
            if (type == 0) fig0(len, ev);
            else if (type == 1) fig1(len, ev);
            else if (type == 2) fig2(len);
            p += len + 1;
        }
    }

    /* text dump used by the parity test (one record per line, fixed field order) */
    std::string dump() const
    {
        std::string o; char t[256];
        snprintf(t, sizeof t, "E %u %u %d %u [%s] [%s]\n", eid, ecc, (int)ensLabel.charset, ensLabel.fig1_flag, hex(ensLabel.fig1_label).c_str(), shortHex(ensLabel).c_str()); o += t;
        o += extDump("XE 0 0", ensLabel);
        for (const auto& s : services) {
            snprintf(t, sizeof t, "S %u %d %d %d %u [%s] [%s]\n", s.serviceId, s.language, s.programType, (int)s.serviceLabel.charset, s.serviceLabel.fig1_flag, hex(s.serviceLabel.fig1_label).c_str(),
                     shortHex(s.serviceLabel).c_str()); o += t;
            snprintf(t, sizeof t, "XS %u 0", s.serviceId); o += extDump(t, s.serviceLabel);
        }
        for (const auto& s : services) for (const auto& c : components) {       /* grouped per service like getComponents() */
            if (c.SId != s.serviceId) continue;
            snprintf(t, sizeof t, "C %u %d %d %d %d %d %u %d %d %d %d %d %u [%s]\n", c.SId, c.componentNr, c.TMid, c.ASCTy, c.DSCTy, c.subchannelId, c.SCId, c.PS_flag, c.CAflag, c.DGflag,
                     c.packetAddress, (int)c.componentLabel.charset, c.componentLabel.fig1_flag, hex(c.componentLabel.fig1_label).c_str()); o += t;
            snprintf(t, sizeof t, "XC %u %d", c.SId, c.componentNr); o += extDump(t, c.componentLabel);
        }
        for (const auto& u : subch) {
            if (u.subChId == -1) continue;
            snprintf(t, sizeof t, "U %d %d %d %d %d %d %d %d %d %d %d\n", u.subChId, u.startAddr, u.length, u.programmeNotData ? 1 : 0, u.protectionSettings.shortForm ? 1 : 0,
                     u.protectionSettings.uepTableIndex, u.protectionSettings.uepLevel, (int)u.protectionSettings.eepProfile, (int)u.protectionSettings.eepLevel, u.language, u.fecScheme); o += t;
        }
        snprintf(t, sizeof t, "T %d %d %d %d %d %d %d %d\n", dateTime.year, dateTime.month, dateTime.day, dateTime.hour, dateTime.minutes, dateTime.seconds, dateTime.hourOffset, dateTime.minuteOffset); o += t;
        return o;
    }

private:
    const uint8_t* fig_ = nullptr; int figAvail_ = 0;

    /* n <= 32 bits starting `bit` bits into the FIG (bit 0 = MSB of the FIG header byte) */
    uint32_t rd(int bit, int n) const
    {
        uint32_t v = 0;
        for (int i = 0; i < n; i++, bit++) {
            const int byte = bit >> 3;
            const int b = byte < figAvail_ ? (fig_[byte] >> (7 - (bit & 7))) & 1 : 0;
            v = (v << 1) | (uint32_t)b;
        }
        return v;
    }
    /* short label, only for labels whose characters are plain ASCII (the EBU Latin -> UTF-8 conversion is not part of this backend) */
    static std::string shortHex(const DabLabel& l) { for (unsigned char c : l.fig1_label) if (c >= 0x7B || c < 0x20 || c == 0x24 || c == 0x5C || c == 0x5E || c == 0x60) return "-"; return hex(l.fig1_shortlabel_utf8()); }
    static std::string hex(const std::string& s) { std::string o; char t[4]; for (unsigned char c : s) { snprintf(t, sizeof t, "%02x", c); o += t; } return o; }
    static std::string extDump(const char* head, const DabLabel& x)
    {
        if (x.segments.empty() && x.segment_count == 0) return std::string();
        std::string o = head; char t[64];
        snprintf(t, sizeof t, " %d %d %d %d", x.toggle_flag ? 1 : 0, (int)x.segment_count, x.fig2_rfu ? 1 : 0, (int)x.extended_label_charset); o += t;
        for (const auto& kv : x.segments) { snprintf(t, sizeof t, " %d:", kv.first); o += t; o += hex(std::string(kv.second.begin(), kv.second.end())); }
        o += " utf8="; o += hex(x.fig2_label());
        return o + "\n";
    }

    void fig0(int len, FigEvents& ev)
    {
        const int pd = (int)rd(8 + 2, 1), ext = (int)rd(8 + 3, 5);
        switch (ext) {
            case 0: {
                const uint16_t e = (uint16_t)rd(16, 16);
                if (e != eid) { eid = e; ev.list.push_back({FigEvents::NewEnsemble, e}); }
                /* CIF counter low part == 0: every twelve seconds in mode I (fib-processor.cpp:132-142, RadioReceiverStats) */
                if (rd(40, 8) % 250 == 0) timeLastFCT0Frame = std::chrono::system_clock::now();
                if (rd(32, 2) != 0) ev.list.push_back({FigEvents::RestartService, 0});
                break;
            }
            case 1: {
                int used = 2;                                       /* bytes */
                while (used < len - 1) {
                    const int o = used * 8;
                    const int id = (int)rd(o, 6);
                    Subchannel& s = subch[id];
                    s.programmeNotData = pd != 0; s.subChId = id; s.startAddr = (int)rd(o + 6, 10);
                    if (rd(o + 16, 1) == 0) {                       /* short form: table index */
                        const int ix = (int)rd(o + 18, 6);
                        s.protectionSettings.uepTableIndex = ix; s.protectionSettings.shortForm = true; s.protectionSettings.uepLevel = kUepProtLevel[ix];
                        s.length = kUepSizeCu[ix];
                        used += 3;
                    } else {
                        s.protectionSettings.shortForm = false;
                        const int option = (int)rd(o + 17, 3);
                        if (option == 0 || option == 1) {
                            s.protectionSettings.eepProfile = option == 0 ? EEPProtectionProfile::EEP_A : EEPProtectionProfile::EEP_B;
                            s.protectionSettings.eepLevel = (EEPProtectionLevel)((int)rd(o + 20, 2) + 1);
                            s.length = (int)rd(o + 22, 10);
                        }
                        used += 4;
                    }
                }
                break;
            }
            case 2: {
                int used = 2;
                while (used < len) {
                    int o = used * 8;
                    uint32_t sid;
                    if (pd) { sid = rd(o, 32); o += 32; } else { sid = rd(o, 16); o += 16; }
                    if (sighting(sid)) ev.list.push_back({FigEvents::ServiceDetected, sid});
                    const int nc = (int)rd(o + 4, 4);
                    o += 8;
                    for (int c = 0; c < nc; c++, o += 16) {
                        const int tmid = (int)rd(o, 2);
                        if (tmid == 2) continue;                    /* reserved */
                        if (!findService(sid) || findComponent(sid, c)) continue;      /* unknown service, or first definition wins */
                        ServiceComponent sc; sc.TMid = (int8_t)tmid; sc.SId = sid; sc.componentNr = (int16_t)c; sc.PS_flag = (int16_t)rd(o + 14, 1);
                        if (tmid == 0) { sc.ASCTy = (int16_t)rd(o + 2, 6); sc.subchannelId = (int16_t)rd(o + 8, 6); }
                        else if (tmid == 1) { sc.DSCTy = (int16_t)rd(o + 2, 6); sc.subchannelId = (int16_t)rd(o + 8, 6); }
                        else { sc.SCId = (uint16_t)rd(o + 2, 12); sc.CAflag = (uint8_t)rd(o + 15, 1); }
                        components.push_back(sc);
                    }
                    used = o / 8;
                }
                break;
            }
            case 3: {
                int used = 2;
                while (used < len) {
                    const int o = used * 8;
                    ServiceComponent* pc = findPacketComponent((int)rd(o, 12));
                    if (pc) { pc->DGflag = (uint8_t)rd(o + 16, 1); pc->DSCTy = (int16_t)rd(o + 18, 6); pc->subchannelId = (int16_t)rd(o + 24, 6); pc->packetAddress = (int16_t)rd(o + 30, 10); }
                    used += 7;
                }
                break;
            }
            case 5: {
                int used = 2;
                while (used < len) {
                    const int o = used * 8;
                    if (rd(o, 1) == 0) { if (rd(o + 1, 1) == 0) subch[rd(o + 2, 6)].language = (int16_t)rd(o + 8, 8); used += 2; }
                    else used += 3;
                }
                break;
            }
            case 9:
                dateTime.hourOffset = rd(16 + 2, 1) ? -(int)rd(16 + 3, 4) : (int)rd(16 + 3, 4);
                dateTime.minuteOffset = rd(16 + 7, 1) ? 30 : 0;
                timeOffsetReceived = true;
                ecc = (uint8_t)rd(16 + 8, 8);
                break;
            case 10: {
                /* Modified Julian Date -> civil date (Fliegel / Van Flandern style integer arithmetic, as fib-processor.cpp:498-517) */
                const int32_t mjd = (int32_t)rd(16 + 1, 17);
                const int32_t j = mjd + 2400001 + 32044;
                const int32_t g = j / 146097, dg = j % 146097;
                const int32_t c = ((dg / 36524) + 1) * 3 / 4, dc = dg - c * 36524;
                const int32_t bb = dc / 1461, db = dc % 1461;
                const int32_t a = ((db / 365) + 1) * 3 / 4, da = db - a * 365;
                const int32_t y = g * 400 + c * 100 + bb * 4 + a;
                const int32_t mo = ((da * 5 + 308) / 153) - 2;
                const int32_t d = da - ((mo + 4) * 153 / 5) + 122;
                dateTime.year = y - 4800 + ((mo + 2) / 12); dateTime.month = ((mo + 2) % 12) + 1; dateTime.day = d + 1;
                dateTime.hour = (int)rd(16 + 21, 5);
                const int minutes = (int)rd(16 + 26, 6);
                if (minutes != dateTime.minutes) dateTime.seconds = 0;
                dateTime.minutes = minutes;
                if (rd(16 + 20, 1) == 1) dateTime.seconds = (int)rd(16 + 32, 6);
                if (timeOffsetReceived) ev.list.push_back({FigEvents::DateTime, 0});
                break;
            }
            case 14: {
                for (int used = 2; used < len; used++) {
                    const int id = (int)rd(used * 8, 6), fec = (int)rd(used * 8 + 6, 2);
                    for (auto& s : subch) if (s.subChId == id) s.fecScheme = (int16_t)fec;
                }
                break;
            }
            case 17: {
                int o = 16;
                while (o < len * 8) {
                    Service* s = findService(rd(o, 16));
                    const bool lflag = rd(o + 18, 1), cc = rd(o + 19, 1);
                    if (lflag) { if (s) s->language = (int16_t)rd(o + 24, 8); o += 8; }
                    if (s) s->programType = (int16_t)rd(o + 27, 5);
                    o += cc ? 40 : 32;
                }
                break;
            }
            default: break;
        }
    }

    void fig1(int len, FigEvents& ev)
    {
        (void)len;
        const int cs = (int)rd(8, 4), oe = (int)rd(8 + 4, 1), ext = (int)rd(8 + 5, 3);
        if (oe) return;
        auto label16 = [&](int bit) { std::string s(16, '\0'); for (int i = 0; i < 16; i++) s[i] = (char)rd(bit + 8 * i, 8); return s; };
        /* the reference keeps the label as a C string: bytes after an embedded NUL are not part of it */
        auto cstr = [](const std::string& s) { return s.substr(0, s.find('\0')); };
        auto set = [&](DabLabel& l, int bit) { const std::string raw = label16(bit); l.fig1_flag = (uint16_t)rd(bit + 128, 16); l.fig1_label = cstr(raw); l.charset = static_cast<CharacterSet>(cs); };
        switch (ext) {
            case 0: if (rd(16, 16) == eid) { set(ensLabel, 32); ev.list.push_back({FigEvents::EnsembleLabel, eid}); } break;
            case 1: { Service* s = findService(rd(16, 16)); if (s) set(s->serviceLabel, 32); break; }
            case 4: {
                const int pd = (int)rd(16, 1), scids = (int)rd(20, 4);
                ServiceComponent* c = findComponent(pd ? rd(24, 32) : rd(24, 16), scids);
                if (c) set(c->componentLabel, pd ? 56 : 40);
                break;
            }
            case 5: { Service* s = findService(rd(16, 32)); if (s) set(s->serviceLabel, 48); break; }
            default: break;
        }
    }

    void extSegment(DabLabel& x, const uint8_t* f, int nbytes, bool toggle, int seg, int rfu)
    {
        if ((x.toggle_flag != 0) != toggle) { x.segments.clear(); x.extended_label_charset = CharacterSet::Undefined; x.toggle_flag = toggle ? 1 : 0; }
        if (seg == 0) {
            x.segment_count = ((f[0] >> 4) & 7) + 1;
            x.extended_label_charset = (f[0] & 0x80) ? CharacterSet::UnicodeUcs2 : CharacterSet::UnicodeUtf8;
            const int skip = rfu == 0 ? 3 : 1;
            if (nbytes <= skip) return;                /* the reference throws here ("FIG2 label length too short"); the glue ignores the FIG */
            f += skip; nbytes -= skip;
            x.fig2_rfu = rfu != 0;
        }
        x.segments[seg] = std::vector<uint8_t>(f, f + nbytes);
    }

    void fig2(int len)
    {
        /* byte view of the FIG like the reference's (always 30 bytes from the FIG start, zero beyond the FIB) */
        uint8_t f[40];                              /* len <= 31: header + 31 bytes, with slack */
        for (int i = 0; i < 40; i++) f[i] = (uint8_t)rd(8 * i, 8);
        const uint8_t* h = f + 1;
        const bool toggle = h[0] & 0x80; const int seg = (h[0] >> 4) & 7, rfu = (h[0] >> 3) & 1, ext = h[0] & 7;
        int idlen;
        switch (ext) { case 0: case 1: idlen = 2; break; case 4: idlen = (h[1] & 0x80) ? 5 : 3; break; case 5: idlen = 4; break; default: return; }
        if (len <= 1 + idlen) return;
        const uint8_t* data = h + 1 + idlen; const int n = len - 1 - idlen;
        if (ext == 0) { if ((uint16_t)(h[1] << 8 | h[2]) == eid) extSegment(ensLabel, data, n, toggle, seg, rfu); }
        else if (ext == 1) { Service* sv = findService(h[1] << 8 | h[2]); if (sv) extSegment(sv->serviceLabel, data, n, toggle, seg, rfu); }
        else if (ext == 4) {
            const int scids = h[1] & 0x0F;
            const uint32_t sid = (h[1] & 0x80) ? ((uint32_t)h[2] << 24 | (uint32_t)h[3] << 16 | (uint32_t)h[4] << 8 | h[5]) : (uint32_t)(h[2] << 8 | h[3]);
            ServiceComponent* c = findComponent(sid, scids);
            if (c) extSegment(c->componentLabel, data, n, toggle, seg, rfu);
        } else {
            const uint32_t sid = (uint32_t)h[1] << 24 | (uint32_t)h[2] << 16 | (uint32_t)h[3] << 8 | h[4];
            Service* sv = findService(sid);
            if (sv) extSegment(sv->serviceLabel, data, n, toggle, seg, rfu);
        }
    }

    /* acceptance rule of fib-processor.cpp:285-327.  When a counter reaches zero the reference calls dropService() with the counter
     * value instead of the SId (:302), i.e. it drops service 0 and the sub-channels no component refers to any more: mirrored. */
    bool sighting(uint32_t sid)
    {
        const auto now = std::chrono::steady_clock::now();
        if (lastDecrement + std::chrono::seconds(1) < now) {
            for (auto it = repeatCount.begin(); it != repeatCount.end();) {
                if (it->second > 0) { it->second--; ++it; }
                else if (it->second == 0) { dropService(0); it = repeatCount.erase(it); }
                else ++it;
            }
            lastDecrement = now;
        }
        int8_t& c = repeatCount[sid];
        if (c < 4) c++;
        if (!findService(sid) && c >= 2) { services.emplace_back(sid); return true; }
        return false;
    }
    void dropService(uint32_t sid)
    {
        for (size_t i = 0; i < services.size();) if (services[i].serviceId == sid) services.erase(services.begin() + i); else i++;
        for (size_t i = 0; i < components.size();) if (components[i].SId == sid) components.erase(components.begin() + i); else i++;
        for (auto& s : subch) {
            if (s.subChId == -1) continue;
            bool used = false;
            for (const auto& c : components) if (c.subchannelId == s.subChId) used = true;
            if (!used) s.subChId = -1;
        }
    }
};

} // namespace dabb_host

#endif
