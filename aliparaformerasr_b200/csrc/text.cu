// Host-side post-processing of the recognisers (SURVEY.md §8 row f2 and the consumer half of f1): tokens.txt table,
// OfflineRecognizer.DecodeMulti / OnlineRecognizer.DecodeMulti token -> text joining, time_stamp_lfr6_onnx.
// Pure C++ on UTF-8 bytes ("▁" = E2 96 81 and "@@" never straddle a code point, so byte-wise IndexOf / Replace give
// the same answers as the reference's UTF-16 string calls).  No device work here.
#include <string.h>

#include <fstream>
#include <string>
#include <vector>

#include "engine.cuh"

namespace pf {

namespace {

const char kBar[] = "\xE2\x96\x81";   // U+2581 LOWER ONE EIGHTH BLOCK, the BPE word-boundary mark

struct TokenTable {
    std::vector<std::string> lines;   // File.ReadAllLines: one entry per line, no terminators
};

// File.ReadAllLines (Utils/PreloadHelper.cs:138): "\r", "\n" and "\r\n" end a line, a final terminator does not open an
// empty last line, a UTF-8 byte-order mark is dropped by the StreamReader.
void split_lines(const char* p, size_t n, std::vector<std::string>& out) {
    size_t i = 0;
    if (n >= 3 && static_cast<unsigned char>(p[0]) == 0xEF && static_cast<unsigned char>(p[1]) == 0xBB && static_cast<unsigned char>(p[2]) == 0xBF) i = 3;
    size_t start = i;
    while (i < n) {
        if (p[i] == '\n' || p[i] == '\r') {
            out.emplace_back(p + start, i - start);
            if (p[i] == '\r' && i + 1 < n && p[i + 1] == '\n') ++i;
            start = ++i;
        } else {
            ++i;
        }
    }
    if (start < n) out.emplace_back(p + start, n - start);
}

// Regex "^[一-龥]+$" (OfflineRecognizer.cs:427-439).  '$' also matches before one final '\n'.
bool is_chinese(const std::string& s) {
    size_t n = s.size();
    if (n && s[n - 1] == '\n') --n;
    if (n == 0) return false;
    size_t i = 0;
    while (i < n) {
        const unsigned char c0 = s[i];
        if ((c0 & 0xF0) != 0xE0 || i + 2 >= n) return false;   // every code point in range is a 3-byte sequence
        const unsigned char c1 = s[i + 1], c2 = s[i + 2];
        if ((c1 & 0xC0) != 0x80 || (c2 & 0xC0) != 0x80) return false;
        const unsigned cp = ((c0 & 0x0Fu) << 12) | ((c1 & 0x3Fu) << 6) | (c2 & 0x3Fu);
        if (cp < 0x4E00 || cp > 0x9FA5) return false;
        i += 3;
    }
    return true;
}

bool is_special(const std::string& s) { return s == "</s>" || s == "<s>" || s == "<blank>" || s == "<unk>"; }

// string.IndexOf(value): first position or -1
long index_of(const std::string& s, const char* needle) {
    const size_t p = s.find(needle);
    return p == std::string::npos ? -1 : static_cast<long>(p);
}

// string.Replace(old, new): left to right, non-overlapping
std::string replace_all(const std::string& s, const std::string& from, const std::string& to) {
    std::string out;
    out.reserve(s.size());
    size_t pos = 0;
    for (;;) {
        const size_t hit = s.find(from, pos);
        if (hit == std::string::npos) break;
        out.append(s, pos, hit - pos);
        out.append(to);
        pos = hit + from.size();
    }
    out.append(s, pos, std::string::npos);
    return out;
}

int count_bars(const std::string& s) {
    int c = 0;
    for (size_t p = s.find(kBar); p != std::string::npos; p = s.find(kBar, p + 3)) ++c;
    return c;
}

// string.Length: UTF-16 code units
int32_t utf16_length(const std::string& s) {
    int32_t n = 0;
    for (unsigned char c : s) {
        if ((c & 0xC0) == 0x80) continue;      // continuation byte
        n += (c >= 0xF0) ? 2 : 1;              // supplementary planes take a surrogate pair
    }
    return n;
}

// string.ToLower() for the scripts a Paraformer vocabulary holds besides CJK: ASCII, Latin-1, Greek, Cyrillic.
// Other code points are copied unchanged.
std::string to_lower(const std::string& s) {
    std::string out;
    out.reserve(s.size());
    size_t i = 0;
    const size_t n = s.size();
    while (i < n) {
        const unsigned char c0 = s[i];
        if (c0 < 0x80) {
            out.push_back(static_cast<char>((c0 >= 'A' && c0 <= 'Z') ? c0 + 32 : c0));
            ++i;
        } else if ((c0 & 0xE0) == 0xC0 && i + 1 < n) {
            unsigned cp = ((c0 & 0x1Fu) << 6) | (static_cast<unsigned char>(s[i + 1]) & 0x3Fu);
            if ((cp >= 0xC0 && cp <= 0xDE && cp != 0xD7) || (cp >= 0x391 && cp <= 0x3AB && cp != 0x3A2)) cp += 32;
            else if (cp >= 0x410 && cp <= 0x42F) cp += 32;
            else if (cp >= 0x400 && cp <= 0x40F) cp += 80;
            out.push_back(static_cast<char>(0xC0 | (cp >> 6)));
            out.push_back(static_cast<char>(0x80 | (cp & 0x3F)));
            i += 2;
        } else {
            const int len = (c0 & 0xF0) == 0xE0 ? 3 : ((c0 & 0xF8) == 0xF0 ? 4 : 1);
            for (int k = 0; k < len && i < n; ++k) out.push_back(s[i++]);
        }
    }
    return out;
}

const std::string& token_at(const TokenTable& t, int32_t id) {
    // _tokens[token] with an id outside the table throws IndexOutOfRangeException in the reference
    if (id < 0 || static_cast<size_t>(id) >= t.lines.size()) throw StatusError{PF_ERR_SHAPE, "token id " + std::to_string(id) + " is outside the tokens table (" + std::to_string(t.lines.size()) + " lines)"};
    return t.lines[id];
}

struct OfflineText {
    std::string text;
    std::vector<std::string> tokens;
    std::vector<std::vector<int32_t>> stamps;
};

// OfflineRecognizer.DecodeMulti, body of the per-stream loop (OfflineRecognizer.cs:310-412)
OfflineText decode_offline(const TokenTable& tab, const int32_t* ids, int n_ids, const int32_t* ts, int n_ts) {
    OfflineText r;
    std::string last_token;
    std::vector<int32_t> last_ts;
    bool have_last_ts = false;
    const std::string bar = kBar;
    const std::string join_mark = std::string("@@") + kBar + kBar;       // "@@▁▁"
    const std::string bar2 = bar + bar, bar3 = bar + bar + bar;
    // List<T>.Remove(list.Last()) removes the FIRST element equal to the last one (:356,:376): for the token strings
    // that can be an earlier duplicate; the int[] timestamps compare by reference, which always resolves to the last.
    auto remove_last_token_by_value = [&](std::vector<std::string>& v) {
        const std::string last = v.back();
        for (size_t i = 0; i < v.size(); ++i)
            if (v[i] == last) {
                v.erase(v.begin() + static_cast<long>(i));
                return;
            }
    };
    const int n = n_ids < n_ts ? n_ids : n_ts;                            // Zip stops at the shorter list (:317)
    for (int i = 0; i < n; ++i) {
        const int32_t id = ids[i];
        if (id == 2) break;
        const std::string& line = token_at(tab, id);
        const std::string cur = line.substr(0, line.find('\t'));           // Split('\t')[0]
        if (is_special(cur)) continue;
        const std::vector<int32_t> stamp = {ts[2 * i], ts[2 * i + 1]};
        if (is_chinese(cur)) {
            r.text += cur;
            r.tokens.push_back(cur);
            r.stamps.push_back(stamp);
            continue;
        }
        r.text += bar + cur + bar;
        const std::string joined = last_token + bar + cur + bar;
        const int bars = count_bars(joined);
        if (index_of(joined, join_mark.c_str()) > 0) {
            const std::string cur_token = replace_all(joined, join_mark, "");
            std::vector<int32_t> cur_ts = stamp;
            if (have_last_ts) {
                cur_ts = last_ts;
                cur_ts.insert(cur_ts.end(), stamp.begin(), stamp.end());
            }
            if (r.tokens.empty() || r.stamps.empty()) throw StatusError{PF_ERR_SHAPE, "DecodeMulti: Last() on an empty list"};
            remove_last_token_by_value(r.tokens);
            r.tokens.push_back(replace_all(cur_token, bar, ""));
            r.stamps.pop_back();
            r.stamps.push_back(cur_ts);
            last_token = cur_token;
            last_ts = cur_ts;
            have_last_ts = true;
        } else if ((bars == 3 || bars == 5) && index_of(joined, bar3.c_str()) < 0) {
            const std::string cur_token = replace_all(joined, bar2, "");
            std::vector<int32_t> cur_ts = stamp;
            if (have_last_ts) {
                cur_ts = last_ts;
                cur_ts.insert(cur_ts.end(), stamp.begin(), stamp.end());
            }
            if (!r.tokens.empty()) remove_last_token_by_value(r.tokens);
            r.tokens.push_back(replace_all(cur_token, bar, ""));
            if (!r.stamps.empty()) r.stamps.pop_back();
            r.stamps.push_back(cur_ts);
            last_token = cur_token;
            last_ts = cur_ts;
            have_last_ts = true;
        } else {
            r.tokens.push_back(replace_all(cur, bar, ""));
            r.stamps.push_back(stamp);
            last_token = bar + cur + bar;
            last_ts = stamp;
            have_last_ts = true;
        }
    }
    std::string& t = r.text;
    if (index_of(t, join_mark.c_str()) > 0 || index_of(t, bar3.c_str()) < 0) {
        t = replace_all(replace_all(replace_all(replace_all(t, join_mark, ""), bar2, " "), "@@", " "), bar, " ");
    } else {
        t = replace_all(replace_all(replace_all(t, bar3, " "), bar2, ""), bar, "");
    }
    return r;
}

// OnlineRecognizer.DecodeMulti, body of the per-stream loop (OnlineRecognizer.cs:408-432): whole lines, no '\t' split
std::string decode_online(const TokenTable& tab, const int32_t* ids, int n_ids) {
    std::string text;
    const std::string bar = kBar;
    for (int i = 0; i < n_ids; ++i) {
        if (ids[i] == 2) break;
        const std::string& cur = token_at(tab, ids[i]);
        if (is_special(cur)) continue;
        text += is_chinese(cur) ? cur : bar + cur + bar;
    }
    text = replace_all(text, std::string("@@") + kBar + kBar, "");
    text = replace_all(text, std::string("@@") + kBar, "");
    text = replace_all(text, bar + bar, " ");
    text = replace_all(text, bar, "");
    return to_lower(text);
}

// OfflineRecognizer.time_stamp_lfr6_onnx (OfflineRecognizer.cs:200-302), float32 arithmetic as the C#.
std::vector<int32_t> timestamps_lfr6(const float* us_cif_peak, int num_frames, const int32_t* tokens, int n_tokens, float begin_time, float total_offset) {
    const int kStartEndThreshold = 5, kMaxTokenDuration = 30;
    const float time_rate = 10.0f * 6 / 1000 / 3;                          // 3 times upsampled
    if (n_tokens < 1) throw StatusError{PF_ERR_SHAPE, "time_stamp_lfr6_onnx: empty token row"};
    if (tokens[n_tokens - 1] == 2) --n_tokens;
    std::vector<float> fire;
    for (int i = 0; i < num_frames; ++i)
        if (static_cast<double>(us_cif_peak[i]) > 1.0 - 1e-4) fire.push_back(static_cast<float>(i) + total_offset);
    if (fire.empty()) throw StatusError{PF_ERR_SHAPE, "time_stamp_lfr6_onnx: us_cif_peak has no fire (fire_place[0] throws in the reference)"};
    std::vector<float> t0, t1;
    std::vector<char> new_char;
    if (fire[0] > kStartEndThreshold) {                                    // begin silence
        t0.push_back(0.0f);
        t1.push_back(fire[0] * time_rate);
        new_char.push_back(0);
    }
    const int nf = static_cast<int>(fire.size());
    for (int i = 0; i < nf - 1; ++i) {
        if (i >= n_tokens) throw StatusError{PF_ERR_SHAPE, "time_stamp_lfr6_onnx: more fires than tokens (tokens[i] throws in the reference)"};
        new_char.push_back(tokens[i] == 1 ? 0 : 1);
        if (i == nf - 2 || fire[i + 1] - fire[i] < kMaxTokenDuration) {
            t0.push_back(fire[i] * time_rate);
            t1.push_back(fire[i + 1] * time_rate);
        } else {
            const float split = fire[i] + kMaxTokenDuration;
            t0.push_back(fire[i] * time_rate);
            t1.push_back(split * time_rate);
            t0.push_back(split * time_rate);
            t1.push_back(fire[i + 1] * time_rate);
            new_char.push_back(0);
        }
    }
    if (t1.empty()) throw StatusError{PF_ERR_SHAPE, "time_stamp_lfr6_onnx: Last() on an empty timestamp list"};
    if (static_cast<float>(num_frames) - fire.back() > kStartEndThreshold) {   // tail token and end silence
        const float end = (static_cast<float>(num_frames) + fire.back()) / 2;
        t1.back() = end * time_rate;
        t0.push_back(end * time_rate);
        t1.push_back(static_cast<float>(num_frames) * time_rate);
        new_char.push_back(0);
    } else {
        t1.back() = static_cast<float>(num_frames) * time_rate;
    }
    if (begin_time > 0.0f) {
        for (size_t i = 0; i < t0.size(); ++i) {
            t0[i] = t0[i] + begin_time / 1000.0f;
            t1[i] = t1[i] + begin_time / 1000.0f;
        }
    }
    new_char.push_back(1);
    std::vector<int32_t> out;
    const size_t n = new_char.size() < t0.size() ? new_char.size() : t0.size();
    for (size_t i = 0; i < n; ++i)
        if (new_char[i]) {
            out.push_back(static_cast<int32_t>(t0[i] * 1000));
            out.push_back(static_cast<int32_t>(t1[i] * 1000));
        }
    return out;
}

template <typename F>
pf_status text_guard(F&& f) {
    try {
        f();
        return PF_OK;
    } catch (const StatusError& e) {
        set_last_error(e.what);
        return e.code;
    } catch (const std::bad_alloc&) {
        set_last_error("host allocation failed");
        return PF_ERR_OOM;
    } catch (const std::exception& e) {
        set_last_error(e.what());
        return PF_ERR_BAD_ARG;
    }
}

void fill_text_result(const OfflineText& r, pf_text_result* out) {
    size_t tok_bytes = 0, ts_count = 0;
    for (const auto& t : r.tokens) tok_bytes += t.size() + 1;
    for (const auto& s : r.stamps) ts_count += s.size();
    out->text_bytes = r.text.size();
    out->text_len = utf16_length(r.text);
    out->n_tokens = static_cast<int32_t>(r.tokens.size());
    out->n_timestamps = static_cast<int32_t>(r.stamps.size());
    out->tokens_bytes = tok_bytes;
    out->ts_count = ts_count;
    const bool text_fits = !out->text || out->text_capacity >= r.text.size() + 1;
    const bool tok_fits = !out->tokens || out->tokens_capacity >= tok_bytes;
    const bool ts_fits = !out->ts || (out->ts_capacity >= ts_count && out->ts_offsets && out->ts_offsets_capacity >= r.stamps.size() + 1);
    if (!text_fits || !tok_fits || !ts_fits) throw StatusError{PF_ERR_BAD_ARG, "pf_text_result buffer too small (required sizes are filled in)"};
    if (out->text) memcpy(out->text, r.text.c_str(), r.text.size() + 1);
    if (out->tokens) {
        char* p = out->tokens;
        for (const auto& t : r.tokens) {
            memcpy(p, t.c_str(), t.size() + 1);
            p += t.size() + 1;
        }
    }
    if (out->ts) {
        size_t k = 0;
        for (size_t i = 0; i < r.stamps.size(); ++i) {
            out->ts_offsets[i] = static_cast<int32_t>(k);
            for (int32_t v : r.stamps[i]) out->ts[k++] = v;
        }
        out->ts_offsets[r.stamps.size()] = static_cast<int32_t>(k);
    }
}

}  // namespace

}  // namespace pf

struct pf_tokens {
    pf::TokenTable table;
};

extern "C" {

pf_status pf_tokens_create_from_memory(const char* utf8, size_t bytes, pf_tokens** out) {
    return pf::text_guard([&] {
        if (!out) throw pf::StatusError{PF_ERR_BAD_ARG, "null out pointer"};
        *out = nullptr;
        if (!utf8 && bytes) throw pf::StatusError{PF_ERR_BAD_ARG, "null tokens text"};
        auto* h = new pf_tokens();
        pf::split_lines(utf8, bytes, h->table.lines);
        if (h->table.lines.empty()) {            // OfflineRecognizer.cs:30: "Tokens file is invalid"
            delete h;
            throw pf::StatusError{PF_ERR_BAD_ARG, "tokens table is empty"};
        }
        *out = h;
    });
}

pf_status pf_tokens_create(const char* path, pf_tokens** out) {
    return pf::text_guard([&] {
        if (!out) throw pf::StatusError{PF_ERR_BAD_ARG, "null out pointer"};
        *out = nullptr;
        if (!path || !*path) throw pf::StatusError{PF_ERR_BAD_ARG, "empty tokens path"};
        std::ifstream f(path, std::ios::binary);
        if (!f) throw pf::StatusError{PF_ERR_BAD_ARG, std::string("cannot open tokens file ") + path};
        std::string data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
        pf_tokens* h = nullptr;
        const pf_status st = pf_tokens_create_from_memory(data.data(), data.size(), &h);
        if (st != PF_OK) throw pf::StatusError{st, pf_last_error()};
        *out = h;
    });
}

pf_status pf_tokens_destroy(pf_tokens* t) {
    if (!t) {
        pf::set_last_error("null tokens handle");
        return PF_ERR_DISPOSED;
    }
    delete t;
    return PF_OK;
}

int32_t pf_tokens_count(const pf_tokens* t) { return t ? static_cast<int32_t>(t->table.lines.size()) : -1; }

int32_t pf_tokens_get(const pf_tokens* t, int32_t id, char* buf, size_t capacity) {
    if (!t || id < 0 || static_cast<size_t>(id) >= t->table.lines.size()) return -1;
    const std::string& s = t->table.lines[id];
    if (buf && capacity) {
        const size_t n = s.size() < capacity - 1 ? s.size() : capacity - 1;
        memcpy(buf, s.data(), n);
        buf[n] = 0;
    }
    return static_cast<int32_t>(s.size());
}

pf_status pf_timestamps_lfr6(const float* us_cif_peak, int32_t num_frames, const int32_t* token_ids, int32_t n_ids,
                             float begin_time, float total_offset, int32_t* out_pairs, int32_t capacity_pairs, int32_t* n_pairs) {
    return pf::text_guard([&] {
        if (!us_cif_peak || !token_ids || !n_pairs || num_frames < 0 || n_ids < 0) throw pf::StatusError{PF_ERR_BAD_ARG, "null or negative argument"};
        const std::vector<int32_t> v = pf::timestamps_lfr6(us_cif_peak, num_frames, token_ids, n_ids, begin_time, total_offset);
        *n_pairs = static_cast<int32_t>(v.size() / 2);
        if (out_pairs) {
            if (capacity_pairs < *n_pairs) throw pf::StatusError{PF_ERR_BAD_ARG, "timestamp buffer too small (n_pairs holds the required count)"};
            if (!v.empty()) memcpy(out_pairs, v.data(), v.size() * sizeof(int32_t));
        }
    });
}

pf_status pf_decode_offline(const pf_tokens* t, const int32_t* token_ids, int32_t n_ids, const int32_t* timestamps,
                            int32_t n_timestamps, pf_text_result* out) {
    return pf::text_guard([&] {
        if (!t) throw pf::StatusError{PF_ERR_DISPOSED, "null tokens handle"};
        if (!out || n_ids < 0 || n_timestamps < 0 || (n_ids && !token_ids)) throw pf::StatusError{PF_ERR_BAD_ARG, "null or negative argument"};
        std::vector<int32_t> zeros;
        if (!timestamps) {                       // 3-output models: one {0, 0} per picked id (OfflineRecognizer.cs:151)
            zeros.assign(2 * static_cast<size_t>(n_ids), 0);
            timestamps = zeros.data();
            n_timestamps = n_ids;
        }
        pf::fill_text_result(pf::decode_offline(t->table, token_ids, n_ids, timestamps, n_timestamps), out);
    });
}

pf_status pf_decode_offline_result(const pf_tokens* t, const pf_result* res, int32_t utt, pf_text_result* out) {
    return pf::text_guard([&] {
        if (!t) throw pf::StatusError{PF_ERR_DISPOSED, "null tokens handle"};
        if (!res || !out || !res->tokens) throw pf::StatusError{PF_ERR_BAD_ARG, "null result"};
        if (utt < 0 || utt >= res->batch) throw pf::StatusError{PF_ERR_BAD_ARG, "utterance index outside the batch"};
        const int32_t L = res->max_len;
        const int32_t* ids = res->tokens + static_cast<size_t>(utt) * L;
        std::vector<int32_t> stamps;
        if (res->us_cif_peak && res->us_frames > 0) {      // cif_peak_tensor != null (OfflineRecognizer.cs:172-183)
            stamps = pf::timestamps_lfr6(res->us_cif_peak + static_cast<size_t>(utt) * res->us_frames, res->us_frames, ids, L, 0.0f, -1.5f);
        } else {
            stamps.assign(2 * static_cast<size_t>(L), 0);
        }
        pf::fill_text_result(pf::decode_offline(t->table, ids, L, stamps.data(), static_cast<int>(stamps.size() / 2)), out);
    });
}

pf_status pf_decode_online(const pf_tokens* t, const int32_t* token_ids, int32_t n_ids, char* text, size_t capacity, size_t* text_bytes) {
    return pf::text_guard([&] {
        if (!t) throw pf::StatusError{PF_ERR_DISPOSED, "null tokens handle"};
        if (n_ids < 0 || (n_ids && !token_ids)) throw pf::StatusError{PF_ERR_BAD_ARG, "null or negative argument"};
        const std::string s = pf::decode_online(t->table, token_ids, n_ids);
        if (text_bytes) *text_bytes = s.size();
        if (text) {
            if (capacity < s.size() + 1) throw pf::StatusError{PF_ERR_BAD_ARG, "text buffer too small (text_bytes holds the required size)"};
            memcpy(text, s.c_str(), s.size() + 1);
        }
    });
}

}  // extern "C"
