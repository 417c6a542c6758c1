// nodes.cpp — see nodes.hpp.
#include "nodes.hpp"

#include <cstdio>

namespace cro {
namespace nodes {

namespace {

const char kRFC3339[] = "2006-01-02T15:04:05Z07:00";

// time.quote (Go stdlib time/format.go): bytes outside printable ASCII as \xNN, '"' and '\' escaped.
std::string tquote(const std::string& s) {
    static const char hex[] = "0123456789abcdef";
    std::string o = "\"";
    for (unsigned char c : s) {
        if (c >= 0x80 || c < ' ') {
            o += "\\x";
            o.push_back(hex[c >> 4]);
            o.push_back(hex[c & 0xF]);
        } else {
            if (c == '"' || c == '\\') o.push_back('\\');
            o.push_back((char)c);
        }
    }
    return o + "\"";
}

bool isDigitAt(const std::string& s, size_t i) { return i < s.size() && s[i] >= '0' && s[i] <= '9'; }

// time.getnum: one or two digits (exactly two when fixed)
bool getnum(std::string* s, bool fixed, int* out) {
    if (!isDigitAt(*s, 0)) return false;
    if (!isDigitAt(*s, 1)) {
        if (fixed) return false;
        *out = (*s)[0] - '0';
        s->erase(0, 1);
        return true;
    }
    *out = ((*s)[0] - '0') * 10 + ((*s)[1] - '0');
    s->erase(0, 2);
    return true;
}

long long daysFromCivil(long long y, unsigned m, unsigned d) {   // proleptic Gregorian, days since 1970-01-01
    y -= m <= 2;
    const long long era = (y >= 0 ? y : y - 399) / 400;
    const unsigned yoe = (unsigned)(y - era * 400);
    const unsigned doy = (153 * (m + (m > 2 ? -3 : 9)) + 2) / 5 + d - 1;
    const unsigned doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
    return era * 146097 + (long long)doe - 719468;
}

int daysIn(int month, int year) {
    static const int n[] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
    if (month == 2 && (year % 4 == 0 && (year % 100 != 0 || year % 400 == 0))) return 29;
    return n[month - 1];
}

}  // namespace

bool ParseRFC3339(const std::string& avalue, long long* unixSeconds, long long* nanos, std::string* err) {
    std::string value = avalue;
    auto cannot = [&](const std::string& valueElem, const std::string& layoutElem) {
        *err = "parsing time " + tquote(avalue) + " as " + tquote(kRFC3339) + ": cannot parse " + tquote(valueElem) + " as " +
               tquote(layoutElem);
        return false;
    };
    auto message = [&](const std::string& m) {
        *err = "parsing time " + tquote(avalue) + m;
        return false;
    };
    auto skip = [&](char lit) {   // time.skip for a one-byte literal prefix
        if (value.empty() || value[0] != lit) return false;
        value.erase(0, 1);
        return true;
    };
    int year = 0, month = 0, day = 0, hour = 0, min = 0, sec = 0;
    long long nsec = 0, zoneOffset = 0;
    std::string hold;

    // "2006"
    hold = value;
    if (value.size() < 4 || !isDigitAt(value, 0)) return cannot(hold, "2006");
    for (int i = 0; i < 4; ++i) {
        if (!isDigitAt(value, (size_t)i)) return cannot(hold, "2006");   // atoi leaves a remainder
        year = year * 10 + (value[(size_t)i] - '0');
    }
    value.erase(0, 4);
    // "-01"
    if (!skip('-')) return cannot(value, "-");
    hold = value;
    if (!getnum(&value, true, &month)) return cannot(hold, "01");
    if (month <= 0 || month > 12) return message(": month out of range");
    // "-02"
    if (!skip('-')) return cannot(value, "-");
    hold = value;
    if (!getnum(&value, true, &day)) return cannot(hold, "02");
    // "T15"
    if (!skip('T')) return cannot(value, "T");
    hold = value;
    if (!getnum(&value, false, &hour)) return cannot(hold, "15");
    if (hour >= 24) return message(": hour out of range");
    // ":04"
    if (!skip(':')) return cannot(value, ":");
    hold = value;
    if (!getnum(&value, true, &min)) return cannot(hold, "04");
    if (min >= 60) return message(": minute out of range");
    // ":05" (+ an undeclared fractional second)
    if (!skip(':')) return cannot(value, ":");
    hold = value;
    if (!getnum(&value, true, &sec)) return cannot(hold, "05");
    if (sec >= 60) return message(": second out of range");
    if (value.size() >= 2 && (value[0] == '.' || value[0] == ',') && isDigitAt(value, 1)) {
        size_t n = 2;
        while (isDigitAt(value, n)) ++n;
        size_t nbytes = n > 10 ? 10 : n;                       // parseNanoseconds keeps 9 digits
        long long ns = 0;
        for (size_t i = 1; i < nbytes; ++i) ns = ns * 10 + (value[i] - '0');
        for (size_t i = nbytes; i < 10; ++i) ns *= 10;
        nsec = ns;
        value.erase(0, n);
    }
    // "Z07:00"
    hold = value;
    if (!value.empty() && value[0] == 'Z') {
        value.erase(0, 1);
    } else {
        if (value.size() < 6 || value[3] != ':') return cannot(hold, "Z07:00");
        const char sign = value[0];
        std::string hh = value.substr(1, 2), mm = value.substr(4, 2);
        value.erase(0, 6);
        int hr = 0, mn = 0;
        bool ok = getnum(&hh, true, &hr) && getnum(&mm, true, &mn);
        std::string range;
        if (hr > 24) range = "time zone offset hour";
        if (mn > 60) range = "time zone offset minute";
        zoneOffset = ((long long)hr * 60 + mn) * 60;
        if (sign == '-') zoneOffset = -zoneOffset;
        else if (sign != '+') ok = false;
        if (!range.empty()) return message(": " + range + " out of range");
        if (!ok) return cannot(hold, "Z07:00");
    }
    if (!value.empty()) return message(": extra text: " + tquote(value));
    if (day < 1 || day > daysIn(month, year)) return message(": day out of range");
    *unixSeconds = daysFromCivil(year, (unsigned)month, (unsigned)day) * 86400 + hour * 3600LL + min * 60LL + sec - zoneOffset;
    *nanos = nsec;
    return true;
}

std::string FormatRFC3339UTC(long long t) {
    long long days = t / 86400, rem = t % 86400;
    if (rem < 0) { rem += 86400; --days; }
    // civil_from_days
    days += 719468;
    const long long era = (days >= 0 ? days : days - 146096) / 146097;
    const unsigned doe = (unsigned)(days - era * 146097);
    const unsigned yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
    long long y = (long long)yoe + era * 400;
    const unsigned doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
    const unsigned mp = (5 * doy + 2) / 153;
    const unsigned d = doy - (153 * mp + 2) / 5 + 1;
    const unsigned m = mp < 10 ? mp + 3 : mp - 9;
    y += m <= 2;
    char buf[40];
    snprintf(buf, sizeof buf, "%04lld-%02u-%02uT%02lld:%02lld:%02lldZ", y, m, d, rem / 3600, rem % 3600 / 60, rem % 60);
    return buf;
}

Error RestartDaemonsetDecision(const std::string& ns, const std::string& name, const DaemonSetView& ds,
                               long long nowUnix, long long nowNanos, Restart* out) {
    *out = Restart::Skipped;
    if (ds.DesiredNumberScheduled == 0) return Error::Nil();                       // :40-43
    if (ds.NumberReady < ds.DesiredNumberScheduled || ds.CurrentNumberScheduled < ds.DesiredNumberScheduled ||
        ds.NumberUnavailable > 0 || ds.NumberMisscheduled > 0)
        return Error::Nil();                                                       // :44-50 not stable: leave it alone
    if (ds.hasRestartedAt) {
        long long last = 0, lastNs = 0;
        std::string perr;
        if (!ParseRFC3339(ds.restartedAt, &last, &lastNs, &perr))
            return Error::New("failed to parse restartedAt annotation for DaemonSet " + ns + "/" + name + ": '" + perr + "'");
        // time.Since(last) <= 10 s, in nanoseconds (a Duration saturates far outside this window)
        const __int128 since = (__int128)(nowUnix - last) * 1000000000 + (nowNanos - lastNs);
        if (since <= (__int128)10 * 1000000000) return Error::Nil();               // :58-62 restarted recently
    }
    *out = Restart::Restarted;                                                     // :69-72 stamp now, Update
    return Error::Nil();
}

}  // namespace nodes
}  // namespace cro
