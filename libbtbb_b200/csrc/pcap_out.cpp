/*
 * pcap_out.cpp -- BR/EDR capture records from batch results: the bytes
 * btbb_pcap_create_file / btbb_pcap_append_packet (pcap.c:74-100, 176-209) write for a packet,
 * serialised from the records the kernels produce (btbb_b200_hit + btbb_b200_decoded) instead
 * of from a btbb_packet.  Pure host formatting, little-endian on the wire as in
 * pcap-common.h:84-97 (DLT 255, LINKTYPE_BLUETOOTH_BREDR_BB).
 *
 * btbb_b200_pcapng_bredr_blocks writes the same packets as pcapng enhanced packet blocks
 * (btbb_pcapng_append_packet, pcapng-bt.c:176-264).
 *
 * For a packet whose payload decode FAILED (rv < 2) the reference still logs the pkt->payload bytes
 * its decoder left behind: ask the decode entry points for records with
 * BTBB_B200_MODE_FLAG_RAW_PAYLOAD and the files are byte-identical for every packet; without the
 * flag such a record comes out with the right length and zero payload bytes.
 */
#include <string.h>
#include <algorithm>
#include <numeric>
#include "../../include/btbb_b200.h"

namespace {

const uint16_t F_DEWHITENED = 0x0001, F_SIGPOWER_VALID = 0x0002, F_NOISEPOWER_VALID = 0x0004, F_REFLAP_VALID = 0x0010,
	       F_PAYLOAD_PRESENT = 0x0020, F_REFUAP_VALID = 0x0080;      /* pcap-common.h:63-76 */
const uint32_t MAX_PAYLOAD = 400;                                     /* BREDR_MAX_PAYLOAD */
const int64_t BB_HEADER = 22;                                         /* pcap_bluetooth_bredr_bb_header without the payload */

inline void le16(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }
inline void le32(uint8_t *p, uint32_t v) { le16(p, v); le16(p + 2, v >> 16); }

}  // namespace

/* the 24-byte file header of btbb_pcap_create_file (pcap.c:49-68, 79-80): nanosecond magic,
 * version 2.4, snaplen 400, DLT 255.  Returns the bytes written, or -1 if cap is too small. */
extern "C" int64_t btbb_b200_pcap_file_header(uint8_t *out, int64_t cap)
{
	if (!out || cap < 24) return -1;
	le32(out, 0xa1b23c4du); le16(out + 4, 2); le16(out + 6, 4);
	le32(out + 8, 0); le32(out + 12, 0); le32(out + 16, MAX_PAYLOAD); le32(out + 20, 255);
	return 24;
}

/* n records of btbb_pcap_append_packet (pcap.c:176-209).  Returns the bytes the records take;
 * they are written only if that fits in cap (out may be NULL to ask for the size). */
extern "C" int64_t btbb_b200_pcap_bredr_records(const btbb_b200_hit *hits, const btbb_b200_decoded *dec,
						const btbb_b200_pcap_meta *meta, int64_t n,
						uint32_t reflap, uint8_t refuap, uint8_t *out, int64_t cap)
{
	if (n < 0 || (n > 0 && (!hits || !dec || !meta))) return -1;
	int64_t need = 0;
	for (int64_t i = 0; i < n; i++) {
		int64_t len = dec[i].payload_length;
		if (len < 0) len = 0;
		if (len > (int64_t)MAX_PAYLOAD) len = MAX_PAYLOAD;
		need += 16 + BB_HEADER + len;
	}
	if (!out || need > cap) return need;
	uint8_t *p = out;
	for (int64_t i = 0; i < n; i++) {
		const btbb_b200_pcap_meta &m = meta[i];
		int64_t caplen = dec[i].payload_length;
		if (caplen < 0) caplen = 0;
		if (caplen > (int64_t)MAX_PAYLOAD) caplen = MAX_PAYLOAD;
		const uint32_t rec = (uint32_t)(BB_HEADER + caplen);
		uint16_t flags = F_DEWHITENED | F_SIGPOWER_VALID;
		if (m.noisedbm < m.sigdbm) flags |= F_NOISEPOWER_VALID;
		if (reflap != BTBB_B200_LAP_ANY) flags |= F_REFLAP_VALID;
		if (refuap != 0xff) flags |= F_REFUAP_VALID;             /* UAP_ANY, btbb.h:96 */
		if (caplen) flags |= F_PAYLOAD_PRESENT;
		le32(p, (uint32_t)(m.ns / 1000000000ull)); le32(p + 4, (uint32_t)(m.ns % 1000000000ull));
		le32(p + 8, rec); le32(p + 12, rec);
		p += 16;
		p[0] = m.channel; p[1] = (uint8_t)m.sigdbm; p[2] = (uint8_t)m.noisedbm; p[3] = hits[i].ac_errors;
		p[4] = (uint8_t)((m.transport << 4) | m.modulation);
		p[5] = 0;                                                   /* corrected header bits: "TODO" upstream */
		le16(p + 6, 0);                                             /* corrected payload bits: likewise */
		le32(p + 8, hits[i].lap);
		le32(p + 12, (reflap & 0xffffffu) | ((uint32_t)refuap << 24));
		le32(p + 16, dec[i].header_packed);
		le16(p + 20, flags);
		p += BB_HEADER;
		/* btbb_get_payload_packed: payload_length bytes; the record carries 344, the rest is zero */
		for (int64_t j = 0; j < caplen; j++) p[j] = j < 344 ? dec[i].payload[j] : 0;
		p += caplen;
	}
	return need;
}

/* n enhanced packet blocks of btbb_pcapng_append_packet (pcapng-bt.c:176-264; block layout
 * pcapng.h, BR/EDR header pcap-common.h:84-97): block type 6, interface 0, the timestamp in
 * nanoseconds split high / low, captured = packet length = 22 + payload bytes, data padded to four
 * bytes (pad bytes are ZERO here; upstream leaves whatever its stack held), a zero "no options" word
 * and the trailing block length.  Same size-query convention as btbb_b200_pcap_bredr_records.  The
 * section header / interface description blocks are the capture program's to write. */
extern "C" int64_t btbb_b200_pcapng_bredr_blocks(const btbb_b200_hit *hits, const btbb_b200_decoded *dec,
						 const btbb_b200_pcap_meta *meta, int64_t n,
						 uint32_t reflap, uint8_t refuap, uint8_t *out, int64_t cap)
{
	if (n < 0 || (n > 0 && (!hits || !dec || !meta))) return -1;
	int64_t need = 0;
	for (int64_t i = 0; i < n; i++) {
		int64_t len = dec[i].payload_length;
		if (len < 0) len = 0;
		if (len > (int64_t)MAX_PAYLOAD) len = MAX_PAYLOAD;
		need += 4 * ((36 + BB_HEADER + len + 3) / 4);
	}
	if (!out || need > cap) return need;
	uint8_t *p = out;
	for (int64_t i = 0; i < n; i++) {
		const btbb_b200_pcap_meta &m = meta[i];
		int64_t caplen = dec[i].payload_length;
		if (caplen < 0) caplen = 0;
		if (caplen > (int64_t)MAX_PAYLOAD) caplen = MAX_PAYLOAD;
		const uint32_t plen = (uint32_t)(BB_HEADER + caplen), blen = 4 * ((36 + plen + 3) / 4);
		uint16_t flags = F_DEWHITENED | F_SIGPOWER_VALID;
		if (m.noisedbm < m.sigdbm) flags |= F_NOISEPOWER_VALID;
		if (reflap != BTBB_B200_LAP_ANY) flags |= F_REFLAP_VALID;
		if (refuap != 0xff) flags |= F_REFUAP_VALID;
		if (caplen) flags |= F_PAYLOAD_PRESENT;
		memset(p, 0, blen);
		le32(p, 6); le32(p + 4, blen); le32(p + 8, 0);
		le32(p + 12, (uint32_t)(m.ns >> 32)); le32(p + 16, (uint32_t)m.ns);
		le32(p + 20, plen); le32(p + 24, plen);
		uint8_t *b = p + 28;
		b[0] = m.channel; b[1] = (uint8_t)m.sigdbm; b[2] = (uint8_t)m.noisedbm; b[3] = hits[i].ac_errors;
		b[4] = (uint8_t)((m.transport << 4) | m.modulation);
		le32(b + 8, hits[i].lap);
		le32(b + 12, (reflap & 0xffffffu) | ((uint32_t)refuap << 24));
		le32(b + 16, dec[i].header_packed);
		le16(b + 20, flags);
		for (int64_t j = 0; j < caplen; j++) b[BB_HEADER + j] = j < 344 ? dec[i].payload[j] : 0;
		le32(p + blen - 4, blen);
		p += blen;
	}
	return need;
}

/* Host helper for the per-piconet entry points (btbb_b200_uap_sieve_*): stable grouping of hit
 * records by LAP, i.e. what get_piconet(LAP) (bluetooth_piconet.c:820-840) does one packet at a
 * time in survey mode.  order[0..n) receives the hit indices sorted by (LAP, arrival order),
 * group_start[0..g] the group boundaries in that order, laps[0..g) each group's LAP (both may
 * be NULL to only count).  Returns the number of groups g (<= n), or -1 on bad arguments. */
extern "C" int64_t btbb_b200_group_by_lap(const btbb_b200_hit *hits, int64_t n, int64_t *order,
					  int64_t *group_start, uint32_t *laps)
{
	if (n < 0 || (n > 0 && (!hits || !order))) return -1;
	std::iota(order, order + n, (int64_t)0);
	std::stable_sort(order, order + n, [&](int64_t a, int64_t b) { return hits[a].lap < hits[b].lap; });
	int64_t g = 0;
	for (int64_t i = 0; i < n; i++)
		if (i == 0 || hits[order[i]].lap != hits[order[i - 1]].lap) {
			if (group_start) group_start[g] = i;
			if (laps) laps[g] = hits[order[i]].lap;
			g++;
		}
	if (group_start) group_start[g] = n;
	return g;
}

